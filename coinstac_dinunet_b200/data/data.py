"""Datasets, data handle (the distributed-training cursor) and the padded sampler.

Parity: coinstac_dinunet/data/data.py:23-242.

B200-first additions
* ``COINNDataHandle.get_loader`` honours ``pin_memory`` and, when the site trains on a
  GPU, wraps the loader in :class:`DevicePrefetcher` (``prefetch_to_device=True`` in
  ``dataloader_args`` or cache) so the H2D copy of batch *t+1* overlaps step *t* on a
  side stream.
* ``COINNPaddedDataSampler`` really shuffles when asked to (reference quirk §8.5-1 keeps
  a fixed order; pass ``reference_order=True`` in the cache to reproduce that).
"""
import json as _json
import math as _math
import os as _os

import numpy as _np
import torch as _torch
from torch.utils.data import DataLoader as _DataLoader, Dataset as _Dataset
from torch.utils.data._utils.collate import default_collate as _default_collate

from .. import config as _conf
from .. import utils as _utils
from ..config.keys import Mode
from ..utils.logger import success
from .datautils import init_k_folds as _init_k_folds

_sep = _os.sep


def safe_collate(batch):
    """``default_collate`` over the items that loaded successfully (falsy items dropped)."""
    return _default_collate([item for item in batch if item])


class COINNDataset(_Dataset):
    """Index-based dataset: ``load_index(file)`` registers samples, ``__getitem__`` loads one."""

    def __init__(self, mode='init', cache=None, input=None, state=None, limit=_conf.max_size):
        self.mode = mode
        self.limit = limit
        self.cache = cache
        self.input = input
        self.state = state
        self.indices = []

    def load_index(self, file):
        """Default: one sample per file.  Override to emit several (e.g. patches)."""
        self.indices.append([file])

    def _load_indices(self, files, **kw):
        for f in files:
            if len(self) >= self.limit:
                break
            self.load_index(f)
        if kw.get('verbose', True):
            print(f'{self.mode}, {len(self)} Indices Loaded')

    def __getitem__(self, index):
        raise NotImplementedError('Must be implemented by child class.')

    def __len__(self):
        return len(self.indices)

    def transforms(self, **kw):
        return None

    def path(self, root_dir='baseDirectory', cache_key='_N/A_'):
        """``state[root_dir] / cache[cache_key]`` - how datasets find their folders."""
        return _os.path.join(self.state[root_dir], self.cache.get(cache_key, ''))

    def add(self, files):
        self._load_indices(files=files, verbose=False)


def _seed_worker(worker_id):
    _np.random.seed((int(_torch.initial_seed()) + worker_id) % (2 ** 32 - 1))


class DevicePrefetcher:
    """Iterate a loader while staging the next ``depth`` batches on ``device`` via a copy stream.

    Batches must be (nested) dict/list/tuple of tensors; tensors are copied with ``non_blocking=True``
    (a true async DMA when the host side is pinned).  Each staged batch carries an event: the consumer
    stream waits on *that* copy only, so later copies keep running under the consumer's kernels.  On
    hand-over every tensor is ``record_stream``-ed on the consumer stream - the caching allocator may
    not recycle the block for a later copy while compute kernels still read it.
    """

    def __init__(self, loader, device, depth=2, stream=None):
        self.loader, self.device = loader, _torch.device(device)
        self.depth = max(int(depth), 1)
        self.dataset = getattr(loader, 'dataset', None)
        self._stream = stream
        if self._stream is None and self.device.type == 'cuda':
            self._stream = _torch.cuda.Stream(self.device)

    def __len__(self):
        return len(self.loader)

    def _walk(self, obj, fn):
        if isinstance(obj, _torch.Tensor):
            return fn(obj)
        if isinstance(obj, dict):
            return {k: self._walk(v, fn) for k, v in obj.items()}
        if isinstance(obj, (list, tuple)):
            return type(obj)(self._walk(v, fn) for v in obj)
        return obj

    def _to_dev(self, obj):
        return self._walk(obj, lambda t: t.to(self.device, non_blocking=True))

    def __iter__(self):
        it = iter(self.loader)
        if self._stream is None:
            yield from it
            return
        from collections import deque
        staged = deque()

        def stage():
            try:
                host = next(it)
            except StopIteration:
                return False
            with _torch.cuda.stream(self._stream):
                dev = self._to_dev(host)
                ev = _torch.cuda.Event()
                ev.record(self._stream)
            staged.append((dev, ev, host))           # `host` stays referenced until its DMA has been consumed
            return True

        more = True
        while more and len(staged) < self.depth:
            more = stage()
        while staged:
            dev, ev, _host = staged.popleft()
            cur = _torch.cuda.current_stream(self.device)
            cur.wait_event(ev)
            self._walk(dev, lambda t: (t.record_stream(cur), t)[1] if t.is_cuda else t)
            if more:
                more = stage()
            yield dev


class COINNDataHandle:
    """Owns the per-site datasets (stored in ``cache['dataset']``) and the train cursor."""

    def __init__(self, cache=None, input=None, state=None, dataloader_args=None, **kw):
        self.cache = cache
        self.input = input
        self.state = state
        self.dataset = self.cache.setdefault('dataset', {})
        args = cache.get('dataloader_args', dataloader_args)
        self.dataloader_args = _utils.FrozenDict(args if args else {})

    # ---------------------------------------------------------------- datasets
    def get_dataset(self, handle_key, files, dataset_cls=None):
        ds = dataset_cls(mode=handle_key, cache=self.cache, input=self.input, state=self.state,
                         limit=self.cache.get('load_limit', _conf.max_size))
        ds.add(files=files)
        self.dataset[handle_key] = ds if len(ds) > 0 else None
        return self.dataset[handle_key]

    def _read_split(self):
        with open(self.cache['split_dir'] + _sep + self.cache['split_file']) as fp:
            return _json.loads(fp.read())

    def _preset(self, handle_key):
        return self.dataloader_args.get(handle_key, {}).get('dataset')

    def get_train_dataset(self, dataset_cls):
        if dataset_cls is None or self._preset('train'):
            return self._preset('train')
        return self.get_dataset('train', self._read_split().get('train', []), dataset_cls=dataset_cls)

    def get_validation_dataset(self, dataset_cls):
        if dataset_cls is None or self._preset('validation'):
            return self._preset('validation')
        ds = self.get_dataset('validation', self._read_split().get('validation', []), dataset_cls=dataset_cls)
        return ds if ds and len(ds) > 0 else None

    def get_test_dataset(self, dataset_cls):
        if dataset_cls is None or self._preset('test'):
            return self._preset('test')
        files = self._read_split().get('test', [])[:self.cache.get('load_limit', _conf.max_size)]
        if self.cache.get('load_sparse') and len(files) > 1:
            # one dataset per subject so predictions can be stitched per subject
            datasets = [self.get_dataset('test', [f], dataset_cls=dataset_cls) for f in files]
            success(f'\n{len(datasets)} sparse dataset loaded.', self.cache.get('verbose'))
            total = sum(len(d) for d in datasets if d)
        else:
            datasets = self.get_dataset('test', files, dataset_cls=dataset_cls)
            total = len(datasets) if datasets else 0
        return datasets if total > 0 else None

    # ----------------------------------------------------------------- loaders
    _LOADER_DEFAULTS = dict(dataset=None, batch_size=1, sampler=None, shuffle=False, batch_sampler=None,
                            num_workers=0, pin_memory=False, drop_last=False, timeout=0)

    def get_loader(self, handle_key='', use_padded_sampler=False, **kw):
        """DataLoader from ``cache`` ∪ ``dataloader_args[handle_key]`` ∪ ``kw`` (later wins)."""
        merged = {**self.cache}
        merged.update(self.dataloader_args.get(handle_key, {}))
        merged.update(kw)

        largs = {k: merged.get(k, dflt) for k, dflt in self._LOADER_DEFAULTS.items()}
        largs['worker_init_fn'] = _seed_worker if merged.get('seed_all') else None
        if largs['num_workers'] and merged.get('persistent_workers'):
            largs['persistent_workers'] = True

        if use_padded_sampler:
            want_shuffle = bool(largs['shuffle']) and not merged.get('reference_order', False)
            largs['sampler'] = COINNPaddedDataSampler(
                largs['dataset'], largs['batch_size'], seed=int(merged.get('seed', 0) or 0),
                shuffle=want_shuffle, drop_last=False)
            largs['sampler'].set_epoch(int(merged.get('epoch', 0) or 0))
            largs['shuffle'] = False
            largs['drop_last'] = False

        loader = _DataLoader(collate_fn=merged.get('collate_fn', safe_collate), **largs)
        dev = merged.get('prefetch_to_device')
        if dev is True:                                   # "the device this site trains on"
            dev = (self.cache.get('device') or {}).get('gpu')
            dev = dev if (dev is not None and _torch.device(dev).type == 'cuda') else None
        if dev:
            dev = _torch.device(dev)
            stream = None
            if dev.type == 'cuda':                        # one copy stream per site, kept across rounds
                streams = self.cache.setdefault('_prefetch_streams', {})
                stream = streams.get(str(dev))
                if stream is None:
                    stream = streams[str(dev)] = _torch.cuda.Stream(dev)
            return DevicePrefetcher(loader, dev, depth=int(merged.get('prefetch_depth', 2) or 2), stream=stream)
        return loader

    def next_iter(self, handle_key=Mode.TRAIN, shuffle=True) -> tuple:
        """Next training batch + control flags.

        ``cursor == 0`` (re)creates the iterator (padded to whole batches).  When the local
        epoch is exhausted the site announces ``mode = validation_waiting`` and rewinds, so
        a lagging site keeps contributing gradients until every site is waiting
        (ref data.py:175-191).
        """
        out = {}
        if self.cache['cursor'] == 0:
            loader = self.get_loader(handle_key=handle_key, shuffle=shuffle,
                                     dataset=self.dataset[handle_key], use_padded_sampler=True,
                                     epoch=self.cache.get('local_epoch', 0))
            self.cache['data_len'] = len(loader) * self.cache['batch_size']
            self.cache['train_loader_iter'] = iter(loader)
            self.cache['local_epoch'] = self.cache.get('local_epoch', 0) + 1

        batch = next(self.cache['train_loader_iter'])
        self.cache['cursor'] += self.cache['batch_size']
        if self.cache['cursor'] >= self.cache['data_len']:
            out['mode'] = Mode.VALIDATION_WAITING
            self.cache['cursor'] = 0
        return batch, out

    # -------------------------------------------------------------------- misc
    def prepare_data(self):
        return _init_k_folds(self.list_files(), self.cache, self.state)

    def list_files(self) -> list:
        if self.cache.get('data_dir'):
            return sorted(_os.listdir(self.state['baseDirectory'] + _sep + self.cache['data_dir']))
        return []


class COINNPaddedDataSampler:
    """Sampler whose length is a whole number of batches (wrap-around padding).

    Every site therefore yields ``ceil(n / batch)`` full batches - a static shape per step,
    which is also what lets the training step be captured in a CUDA graph.
    """

    def __init__(self, dataset, batch_size, seed=0, shuffle=False, drop_last=False):
        self.dataset = dataset
        self.batch_size = int(batch_size)
        self.drop_last = drop_last
        self.shuffle = shuffle
        self.seed = seed
        self.epoch = 0
        n_batches = len(dataset) / self.batch_size
        self.total_size = int((_math.floor if drop_last else _math.ceil)(n_batches)) * self.batch_size

    def __iter__(self):
        n = len(self.dataset)
        if self.shuffle:
            gen = _torch.Generator()
            gen.manual_seed(self.seed + self.epoch)
            order = _torch.randperm(n, generator=gen).tolist()
        else:
            order = list(range(n))
        if self.drop_last:
            order = order[:self.total_size]
        elif n > 0:
            reps = _math.ceil(self.total_size / n)
            order = (order * reps)[:self.total_size]
        assert len(order) == self.total_size
        return iter(order)

    def set_epoch(self, epoch):
        self.epoch = epoch

    def __len__(self):
        return int(self.total_size)
