"""``NNTrainer`` - the per-site training runtime (model/optimizer/device life-cycle,
evaluation loop, purely-local training loop, checkpoint I/O).

API parity: coinstac_dinunet/nn/basetrainer.py:20-326 (hook names, services, cache keys,
checkpoint layout ``{'source','models','optimizers'}`` - SURVEY §5.4).

B200-first differences
* Devices: one site == one GPU.  ``cache['gpus']`` picks it; >=2 ids keep the reference's
  ``nn.DataParallel`` behaviour for drop-in compatibility.
* ``cache['compute_dtype']`` (``'bf16'``/``'fp16'``/``'fp32'``) and ``cache['channels_last']``
  are applied when the model is placed on the device; hand-written sm_100a kernels are swapped
  in by ``ops.nativize`` when ``cache['native_ops']`` is set.
* ``save_checkpoint`` keeps *every* model/optimizer (the reference keeps only the last one,
  quirk §8.5-3); files written by the reference still load.
* No ``.item()`` in the loops: scores are reduced on the device and read when logged.
"""
from collections import OrderedDict as _ODict

import torch as _torch

from .. import config as _conf
from .. import metrics as _base_metrics
from .. import utils as _utils
from ..config.keys import Key, Mode
from ..utils import tensorutils as _tu
from ..utils.logger import info, lazy_debug
from ..utils.utils import stop_training_
from ..vision import plotter as _plot

_DTYPES = {'bf16': _torch.bfloat16, 'bfloat16': _torch.bfloat16, 'fp16': _torch.float16,
           'float16': _torch.float16, 'fp32': _torch.float32, 'float32': _torch.float32, None: None}


def _unwrap(module):
    """Strip DataParallel / DADParallel style wrappers that expose ``.module``."""
    inner = getattr(module, 'module', None)
    return inner if isinstance(inner, _torch.nn.Module) else module


class NNTrainer:
    def __init__(self, data_handle=None, **kw):
        self.cache = data_handle.cache
        self.input = _utils.FrozenDict(data_handle.input)
        self.state = _utils.FrozenDict(data_handle.state)
        self.nn = _ODict()
        self.device = _ODict()
        self.optimizer = _ODict()
        self.data_handle = data_handle

    # ------------------------------------------------------------------ hooks
    def _init_nn_model(self):
        """User hook: fill ``self.nn[name] = nn.Module``."""
        raise NotImplementedError('Must be implemented in child class.')

    def _init_nn_weights(self, **kw):
        """Pretrained checkpoint if given, else (train mode) seeded Kaiming init so every
        site starts from identical weights."""
        if self.cache.get('pretrained_path') is not None:
            self.load_checkpoint(self.cache['pretrained_path'])
        elif self.cache['mode'] == Mode.TRAIN:
            _torch.manual_seed(self.cache['seed'])
            for name in self.nn:
                _tu.initialize_weights(self.nn[name])

    def _init_optimizer(self):
        """Default: Adam(lr=cache['learning_rate']) over the first model."""
        first = next(iter(self.nn))
        self.optimizer['adam'] = _torch.optim.Adam(self.nn[first].parameters(), lr=self.cache['learning_rate'])

    def init_nn(self, init_model=False, init_optim=False, set_devices=False, init_weights=False):
        if init_model:
            self._init_nn_model()
        if init_optim:
            self._init_optimizer()
        if init_weights:
            self._init_nn_weights(init_weights=init_weights)
        if set_devices:
            self._set_gpus()

    def _set_gpus(self):
        self.device['gpu'] = _torch.device('cpu')
        gpus = self.cache.get('gpus')
        if gpus and _torch.cuda.is_available():
            self.device['gpu'] = _torch.device(f'cuda:{gpus[0]}')
            if len(gpus) >= 2:
                for name in self.nn:
                    if not isinstance(self.nn[name], _torch.nn.DataParallel):
                        self.nn[name] = _torch.nn.DataParallel(self.nn[name], gpus)
        for name in self.nn:
            self.nn[name] = self.nn[name].to(self.device['gpu'])
            if self.device['gpu'].type == 'cuda' and self.cache.get('channels_last'):
                self.nn[name] = self.nn[name].to(memory_format=_torch.channels_last_3d
                                                 if self.cache['channels_last'] == '3d'
                                                 else _torch.channels_last)
            if self.device['gpu'].type == 'cuda' and self.cache.get('native_ops') \
                    and not isinstance(self.nn[name], _torch.nn.DataParallel):
                from ..ops.nativize import nativize      # user-defined modules: matching layer patterns -> sm_100a kernels
                self.nn[name] = nativize(self.nn[name])

    @property
    def compute_dtype(self):
        return _DTYPES.get(self.cache.get('compute_dtype'))

    # ------------------------------------------------------------- checkpoints
    def load_checkpoint(self, file_path):
        try:
            chk = _torch.load(file_path, weights_only=False)
        except Exception:
            chk = _torch.load(file_path, map_location='cpu', weights_only=False)

        if isinstance(chk, dict) and str(chk.get('source', 'Unknown')).lower() == 'coinstac':
            for name, sd in chk.get('models', {}).items():
                if name in self.nn:
                    _unwrap(self.nn[name]).load_state_dict(sd)
            for name, sd in chk.get('optimizers', {}).items():
                if name in self.optimizer:
                    self.optimizer[name].load_state_dict(sd)
        else:  # foreign checkpoint: a bare state_dict for the first model
            _unwrap(self.nn[next(iter(self.nn))]).load_state_dict(chk)
        arena = self.cache.get('_arena')
        if arena is not None:  # keep fp32 masters of the fused optimizer in sync
            arena.refresh_from_params()

    def save_checkpoint(self, file_path, src='coinstac'):
        chk = {'source': src, 'models': {}, 'optimizers': {}}
        for name in self.nn:
            chk['models'][name] = _unwrap(self.nn[name]).state_dict()
        for name in self.optimizer:
            chk['optimizers'][name] = self.optimizer[name].state_dict()
        _torch.save(chk, file_path)

    # --------------------------------------------------------------- evaluation
    def evaluation(self, mode='eval', dataset_list=None, save_pred=False, use_padded_sampler=False):
        """Score ``dataset_list`` without gradients; returns ``(averages, metrics)``."""
        for name in self.nn:
            self.nn[name].eval()

        total_avg, total_metrics = self.new_averages(), self.new_metrics()
        loaders = [
            self.data_handle.get_loader(handle_key=mode, dataset=d, shuffle=False,
                                        use_padded_sampler=use_padded_sampler)
            for d in (dataset_list or []) if d is not None and len(d) > 0
        ]

        def fold_in(extra, it, avg, met):
            extra = extra or {}
            avg.accumulate(extra.get('averages', it['averages']))
            met.accumulate(extra.get('metrics', it['metrics']))

        sparse = bool(self.cache.get('load_sparse'))
        verbose = self.cache.get('verbose')
        with _torch.no_grad():
            for loader in loaders:
                avg, met, kept = self.new_averages(), self.new_metrics(), []
                for i, batch in enumerate(loader, 1):
                    it = self.iteration(batch)
                    if save_pred and sparse:
                        kept.append(it)
                    elif save_pred:
                        fold_in(self.save_predictions(loader.dataset, it), it, avg, met)
                    else:
                        fold_in(None, it, avg, met)
                    if verbose and len(loaders) <= 1 and lazy_debug(i):
                        info(f" Itr:{i}/{len(loader)}, Averages:{it.get('averages').get()}, "
                             f"Metrics:{it.get('metrics').get()}")
                if save_pred and sparse and kept:
                    merged = self.reduce_iteration(kept)
                    fold_in(self.save_predictions(loader.dataset, merged), merged, avg, met)
                if verbose and len(loaders) > 1:
                    info(f" {mode}, {avg.get()}, {met.get()}")
                total_metrics.accumulate(met)
                total_avg.accumulate(avg)
        info(f"{mode} metrics: {total_avg.get()}, {total_metrics.get()}", verbose)
        return total_avg, total_metrics

    # ----------------------------------------------------------- local training
    def training_iteration_local(self, i, batch):
        """One micro-batch: forward, backward; optimizer step every ``local_iterations``."""
        it = self.iteration(batch)
        it['loss'].backward()
        if i % self.cache.get('local_iterations', 1) == 0:
            opt = self.optimizer[next(iter(self.optimizer))]
            opt.step()
            opt.zero_grad()
        return it

    def init_training_cache(self):
        self.cache[Key.TRAIN_LOG] = []
        self.cache[Key.VALIDATION_LOG] = []
        self.cache['best_val_epoch'] = 0
        maximize = self.cache['metric_direction'] == 'maximize'
        self.cache['best_val_score'] = 0.0 if maximize else _conf.max_size

    def train_local(self, train_dataset, val_dataset):
        """Single-site training with validation, best-checkpointing and early stop
        (used for pre-training and by ``SiteRunner``; ref basetrainer.py:192-243)."""
        out = {}
        val_list = val_dataset if isinstance(val_dataset, list) else [val_dataset]
        loader = self.data_handle.get_loader('train', dataset=train_dataset, drop_last=True, shuffle=True)
        k = self.cache.get('local_iterations', 1)
        steps = len(loader) // k
        verbose = self.cache.get('verbose')
        ep = 0
        for ep in range(1, self.cache['epochs'] + 1):
            for name in self.nn:
                self.nn[name].train()

            win_avg, win_met = self.new_averages(), self.new_metrics()   # since last log line
            ep_avg, ep_met, micro = self.new_averages(), self.new_metrics(), []
            for i, batch in enumerate(loader, 1):
                micro.append(self.training_iteration_local(i, batch))
                if i % k:
                    continue
                it, micro, step = self.reduce_iteration(micro), [], i // k
                for a, m in ((ep_avg, ep_met), (win_avg, win_met)):
                    a.accumulate(it['averages'])
                    m.accumulate(it['metrics'])
                if lazy_debug(step) or step == steps:
                    info(f"Ep:{ep}/{self.cache['epochs']},Itr:{step}/{steps},{win_avg.get()},{win_met.get()}",
                         verbose)
                    self.cache[Key.TRAIN_LOG].append([*win_avg.get(), *win_met.get()])
                    win_avg.reset(), win_met.reset()
                self.on_iteration_end(i=step, ep=ep, it=it)

            if any(v is not None for v in val_list) and ep % self.cache.get('validation_epochs', 1) == 0:
                info('--- Validation ---', verbose)
                val_avg, val_met = self.evaluation(mode='validation', dataset_list=val_list,
                                                   use_padded_sampler=True)
                self.cache[Key.VALIDATION_LOG].append([*val_avg.get(), *val_met.get()])
                out.update(**self._save_if_better(ep, val_met))
                self._on_epoch_end(ep=ep, ep_averages=ep_avg, ep_metrics=ep_met,
                                   val_averages=val_avg, val_metrics=val_met)
                if lazy_debug(ep):
                    self._save_progress(self.cache, epoch=ep)
                if self._stop_early(ep, val_met, val_averages=val_avg,
                                    epoch_averages=ep_avg, epoch_metrics=ep_met):
                    break

        self._save_progress(self.cache, epoch=ep)
        _utils.save_cache(self.cache, self.cache['log_dir'])
        return out

    # ------------------------------------------------------------- user hooks
    def iteration(self, batch):
        """User hook: one mini-batch.  Must return a dict with at least ``loss`` (a tensor
        to call ``backward`` on), ``averages`` (COINNAverages) and ``metrics`` (COINNMetrics)."""
        return {}

    def save_predictions(self, dataset, its):
        pass

    def reduce_iteration(self, its):
        """Merge micro-batch outputs: averages/metrics accumulate; anything else becomes a
        zero-arg closure producing the concatenated tensor / list on demand (lazy)."""
        merged = {}
        for key, first in its[0].items():
            if isinstance(first, _base_metrics.COINNAverages):
                acc = self.new_averages()
                for it in its:
                    acc.accumulate(it[key])
                merged[key] = acc
            elif isinstance(first, _base_metrics.COINNMetrics):
                acc = self.new_metrics()
                for it in its:
                    acc.accumulate(it[key])
                merged[key] = acc
            else:
                merged[key] = _LazyCollect(key, its)
        return merged

    def _save_if_better(self, epoch, val_metrics):
        return {}

    def new_metrics(self):
        return _base_metrics.COINNMetrics()

    def new_averages(self):
        return _base_metrics.COINNAverages(num_averages=1)

    def _on_epoch_end(self, ep, **kw):
        return {}

    def on_iteration_end(self, i, ep, it):
        return {}

    def _save_progress(self, cache, epoch):
        _plot.plot_progress(cache, self.cache['log_dir'], plot_keys=[Key.TRAIN_LOG], epoch=epoch)
        _plot.plot_progress(cache, self.cache['log_dir'], plot_keys=[Key.VALIDATION_LOG],
                            epoch=epoch // max(self.cache.get('validation_epochs', 1), 1))

    def _stop_early(self, epoch, val_metrics=None, **kw):
        return stop_training_(epoch, self.cache)


class _LazyCollect:
    """Callable standing in for a per-iteration output: concatenates leaf no-grad tensors,
    otherwise returns the list of values (ref basetrainer.py:284-292)."""

    __slots__ = ('key', 'src')

    def __init__(self, key, src):
        self.key, self.src = key, src

    def __call__(self):
        vals = [it[self.key] for it in self.src]
        v0 = vals[0]
        if isinstance(v0, _torch.Tensor) and not v0.requires_grad and v0.is_leaf:
            return _torch.cat([v if v.dim() > 0 else v.unsqueeze(0) for v in vals])
        return vals
