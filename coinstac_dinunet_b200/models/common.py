"""Shared pieces of the model zoo: a classification trainer whose ``iteration`` never syncs
with the host, and in-memory synthetic datasets (there is no network on the build/bench boxes,
so datasets are generated with the documented shapes - BASELINE.json ``data: synthetic``).
"""
import json as _json
import os as _os

import numpy as _np
import torch as _torch

from .. import ops as _ops
from ..data import COINNDataset
from ..trainer import COINNTrainer


class ClassificationTrainer(COINNTrainer):
    """``iteration`` = forward -> fused log-softmax+NLL(+argmax) -> device-side score update.

    Matches the README recipe of the reference (README.md:73-84) except that neither the loss
    (``loss.item()``) nor the metrics (4 ``.item()``) are read back per step.
    Sub-classes define ``_init_nn_model`` and may override ``_inputs``.
    """

    input_key, label_key = 'inputs', 'labels'

    def _inputs(self, batch):
        dev = self.device['gpu']
        x = batch[self.input_key].to(dev, non_blocking=True)
        y = batch[self.label_key].to(dev, non_blocking=True).long()
        dt = self.compute_dtype
        native = getattr(self.nn[next(iter(self.nn))], 'is_native', False)
        if native and x.is_floating_point():
            pass                                         # native kernels read fp32/bf16 inputs directly
        elif dt is not None and x.dtype != dt and x.is_floating_point() and dev.type == 'cuda':
            x = x.to(dt)
        elif not x.is_floating_point() or (dev.type == 'cpu' and x.dtype != _torch.float32):
            x = x.float()
        return x, y

    def _forward(self, model, x):
        """bf16/fp16 compute for stock ``nn.Module``s goes through autocast; native modules (which
        carry their own low-precision shadows) are called directly."""
        dt = self.compute_dtype
        if dt in (_torch.bfloat16, _torch.float16) and x.is_cuda and not getattr(model, 'is_native', False):
            with _torch.autocast('cuda', dtype=dt):
                return model(x)
        return model(x)

    def iteration(self, batch):
        x, y = self._inputs(batch)
        model = self.nn[next(iter(self.nn))]
        logits = self._forward(model, x)
        loss, pred = _ops.softmax_nll(logits, y)
        avg, met = self.new_averages(), self.new_metrics()
        avg.add(loss.detach(), len(y))
        if self.cache.get('monitor_metric') == 'auc':
            met.add(_torch.softmax(logits.detach().float(), 1)[:, 1], y)
        else:
            met.add(pred, y)
        return {'loss': loss, 'averages': avg, 'metrics': met, 'prediction': pred}


def pinned_collate(batch):
    """Collate for in-memory datasets: when the samples of a batch are consecutive views of one
    (pinned) host tensor the batch is a zero-copy slice of it - the H2D DMA then reads pinned
    memory directly; otherwise fall back to stacking."""
    batch = [b for b in batch if b]
    first = batch[0]
    out = {}
    for key in first:
        vals = [b[key] for b in batch]
        v0 = vals[0]
        if isinstance(v0, _torch.Tensor) and v0.dim() > 0 and v0.is_contiguous():
            step = v0.numel() * v0.element_size()
            base = getattr(v0, '_base', None)
            consecutive = base is not None and all(
                getattr(v, '_base', None) is base and v.data_ptr() == v0.data_ptr() + i * step
                for i, v in enumerate(vals))
            if consecutive:
                start = (v0.data_ptr() - base.data_ptr()) // step
                out[key] = base[start:start + len(vals)]
                continue
        out[key] = _torch.stack(vals) if isinstance(v0, _torch.Tensor) else _torch.as_tensor(vals)
    return out


class ArrayFileDataset(COINNDataset):
    """One ``.npy`` feature array per subject in ``<baseDirectory>/<data_dir>`` plus a
    ``labels.json`` (``{file: class}``) next to the folder (``cache['labels_file']``)."""

    def __init__(self, **kw):
        super().__init__(**kw)
        self._labels = None

    def _label_of(self, file):
        if self._labels is None:
            path = self.path(cache_key='labels_file') if self.cache.get('labels_file') else \
                _os.path.join(self.state['baseDirectory'], 'labels.json')
            with open(path) as fp:
                self._labels = _json.load(fp)
        return int(self._labels[file])

    def __getitem__(self, ix):
        file = self.indices[ix][0]
        arr = _np.load(_os.path.join(self.path(cache_key='data_dir'), file))
        return {'inputs': _torch.from_numpy(_np.ascontiguousarray(arr)).float(),
                'labels': _torch.tensor(self._label_of(file), dtype=_torch.long)}


def write_synthetic_site(base_dir, n_subjects, shape, num_class=2, seed=0, data_dir='data', signal=1.0,
                         dtype=_np.float32):
    """Create ``<base_dir>/<data_dir>/subj_XXXX.npy`` + ``labels.json`` with a learnable signal:
    class ``c`` shifts a fixed random direction by ``±signal``."""
    rng = _np.random.default_rng(seed)
    folder = _os.path.join(base_dir, data_dir)
    _os.makedirs(folder, exist_ok=True)
    direction = _np.random.default_rng(12345).standard_normal(shape).astype(_np.float32)
    direction /= _np.sqrt((direction ** 2).mean())
    labels = {}
    for i in range(n_subjects):
        y = int(rng.integers(0, num_class))
        x = rng.standard_normal(shape).astype(_np.float32) + signal * (2.0 * y / max(num_class - 1, 1) - 1.0) * direction
        name = f'subj_{seed:03d}_{i:05d}.npy'
        _np.save(_os.path.join(folder, name), x.astype(dtype))
        labels[name] = y
    with open(_os.path.join(base_dir, 'labels.json'), 'w') as fp:
        _json.dump(labels, fp)
    return folder


class InMemorySynthetic(COINNDataset):
    """Synthetic samples generated once and kept in (optionally pinned) host memory - the data
    source of ``bench.py``.  ``files`` are just integer-like ids."""

    def __init__(self, shape=(66,), num_class=2, n=None, seed=0, pin=False, dtype=_torch.float32, **kw):
        super().__init__(**kw)
        self.shape, self.num_class, self.seed, self.pin, self.dtype = tuple(shape), num_class, seed, pin, dtype
        self._x = self._y = None
        if n:
            self.add([str(i) for i in range(n)])

    def _materialise(self):
        g = _torch.Generator().manual_seed(self.seed)
        n = len(self.indices)
        self._y = _torch.randint(0, self.num_class, (n,), generator=g)
        distinct = min(n, int(self.cache.get('synthetic_distinct', n)) if self.cache else n)
        x = _torch.randn((distinct, *self.shape), generator=g)
        x += (self._y[:distinct].float() * 2 - 1).view(-1, *([1] * len(self.shape))) * 0.5
        self._distinct = distinct   # samples index the pool modulo `distinct` (bounded host memory)
        self._y = self._y[:distinct].contiguous()
        self._x = x.to(self.dtype).contiguous()
        if self._x._base is not None:
            self._x = self._x.clone()
        dev = self.cache.get('synthetic_device') if self.cache else None
        if dev:
            self._x, self._y = self._x.to(dev), self._y.to(dev)
        elif self.pin and _torch.cuda.is_available():
            self._x, self._y = self._x.pin_memory(), self._y.pin_memory()

    def load_index(self, file):
        self.indices.append([file])

    def __getitem__(self, ix):
        if self._x is None:
            self._materialise()
        j = ix % self._distinct
        return {'inputs': self._x[j], 'labels': self._y[j]}
