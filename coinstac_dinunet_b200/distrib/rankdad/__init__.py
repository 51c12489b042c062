"""rankDAD engine: exchange per-layer low-rank (Δ, A) factors instead of dense gradients.

Parity: coinstac_dinunet/distrib/rankdad/__init__.py:12-98 (one round trip per step, no
gradient accumulation, concat across sites along the rank axis then re-compression to
``dad_reduction_rank``).  As in the reference the reconstructed gradient is the *sum* over
sites (no 1/S); set ``cache['dad_mean'] = True`` to divide by the site count.
Unlike the reference the re-compression also runs on CPU aggregators and parameters outside
DAD layers are averaged densely.
"""
import os as _os

import numpy as _np
import torch as _torch

from ...utils import tensorutils as _tu
from ..learner import COINNLearner as _COINNLearner
from ..reducer import COINNReducer as _COINNReducer
from .spi import DADParallel, power_iteration_BC  # noqa: F401


class DADLearner(_COINNLearner):
    def __init__(self, **kw):
        super().__init__(**kw)
        for key in list(self.trainer.nn):
            if not isinstance(self.trainer.nn[key], DADParallel):
                self.trainer.nn[key] = DADParallel(
                    self.trainer.nn[key], cache=self.cache, input=self.input, state=self.state,
                    device=self.trainer.device['gpu'], dtype=self.dtype)
            else:  # re-bind the per-round input/state on the persistent wrapper
                self.trainer.nn[key].input, self.trainer.nn[key].state = self.input, self.state

    def step(self):
        self.model.synced_param_update()
        self.optim.step()
        return {}

    def forward(self):
        """One micro-batch forward + backward (rankDAD cannot accumulate gradients)."""
        out = {}
        self.model.train()
        self.optim.zero_grad()
        batch, flags = self.trainer.data_handle.next_iter()
        it = self.trainer.iteration(batch)
        it['loss'].backward()
        out.update(**flags)
        return self.trainer.reduce_iteration([it]), out

    def to_reduce(self):
        self.model.train()
        it, out = self.forward()
        out.update(**self.model.dad_backward())
        out['reduce'] = True
        return it, out


class DADReducer(_COINNReducer):
    def __init__(self, **kw):
        super().__init__(**kw)
        self.rank = self.cache.setdefault('dad_reduction_rank', 10)
        self.num_pow_iters = self.cache.setdefault('dad_num_pow_iters', 5)
        self.dad_tol = self.cache.setdefault('dad_tol', 1e-3)
        self.recompress = self.cache.setdefault('dad_recompress', True)

    def reduce(self):
        out = {'reduced_dad_data': 'reduced_dad_data.npy'}
        per_site = self._load('dad_data')
        n_sites = len(per_site)
        scale = 1.0 / n_sites if self.cache.get('dad_mean') else 1.0
        reduced = []
        for layer in zip(*per_site):  # same layer from every site
            deltas = [_torch.from_numpy(_np.asarray(p[0], dtype=_np.float32)).to(self.device) for p in layer]
            acts = [_torch.from_numpy(_np.asarray(p[1], dtype=_np.float32)).to(self.device) for p in layer]
            delta = _torch.cat([d.reshape(d.shape[0], -1) for d in deltas], dim=1) * scale   # [out, S·k]
            act = _torch.cat(acts, dim=1)                                                     # [in, S·k]
            if self.recompress and delta.shape[1] > self.rank:
                delta, act = power_iteration_BC(delta, act, self.rank, self.num_pow_iters, self.dad_tol)
            reduced.append([delta.cpu().numpy().astype(self.dtype), act.cpu().numpy().astype(self.dtype)])

        box = _np.empty(len(reduced), dtype=object)
        for i, pair in enumerate(reduced):
            inner = _np.empty(2, dtype=object)
            inner[0], inner[1] = pair
            box[i] = inner
        _tu.save_arrays(self.state['transferDirectory'] + _os.sep + out['reduced_dad_data'], box)

        site0 = next(iter(self.input.values()))
        if site0.get('dad_plain_grads'):
            out['reduced_dad_plain_grads'] = 'reduced_dad_plain_grads.npy'
            dense = self._average('dad_plain_grads')
            if not self.cache.get('dad_mean'):
                dense = [d * n_sites for d in dense]  # keep the same (sum) scaling as the factors
            self._ship(out['reduced_dad_plain_grads'], dense)
        out['update'] = True
        return out
