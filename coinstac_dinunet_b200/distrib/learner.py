"""Site-side half of an aggregation engine: run backward, ship the payload, apply the result.

``COINNLearner`` speaks the reference's *file* wire format (``grads.npy`` up,
``avg_grads.npy`` down; coinstac_dinunet/distrib/learner.py:9-59) so CPU sites and real
COINSTAC deployments keep working.  The NVLink data plane lives in
``coinstac_dinunet_b200.parallel.nvlink_learner`` and subclasses this class, overriding
``to_reduce``/``step`` with the fused sm_100a kernel.
"""
from os import sep as _sep

import numpy as _np
import torch as _torch

from .. import config as _conf
from ..utils import tensorutils as _tu


class COINNLearner:
    def __init__(self, trainer=None, mp_pool=None, **kw):
        self.trainer = trainer
        self.cache = trainer.cache
        self.input = trainer.input
        self.state = trainer.state
        self.pool = mp_pool
        self.global_modes = self.input.get('global_modes', {})
        self.dtype = f"float{self.cache.get('precision_bits', 32)}"
        self.device = trainer.device['gpu']

    # Only the first model / optimizer take part in distributed learning (quirk §8.5-6).
    @property
    def model(self):
        return self.trainer.nn[next(iter(self.trainer.nn))]

    @property
    def optim(self):
        return self.trainer.optimizer[next(iter(self.trainer.optimizer))]

    def _assign_grads(self, arrays):
        """Install averaged gradients with ONE host->device copy of a flat fp32 staging
        buffer, then per-parameter views (the reference issues one copy per parameter)."""
        params = list(self.model.parameters())
        if len(arrays) != len(params):
            raise ValueError(f'{len(arrays)} gradient arrays for {len(params)} parameters')
        flat = _np.concatenate([_np.asarray(a, dtype=_np.float32).reshape(-1) for a in arrays]) \
            if len(arrays) else _np.zeros(0, _np.float32)
        dev_flat = _torch.from_numpy(flat).to(self.device, non_blocking=True)
        off = 0
        for p in params:
            n = p.numel()
            g = dev_flat[off:off + n].view_as(p)
            p.grad = g if g.dtype == p.dtype else g.to(p.dtype)
            off += n

    def step(self) -> dict:
        grads = _tu.load_arrays(self.state['baseDirectory'] + _sep + self.input['avg_grads_file'])
        self._assign_grads(list(grads))
        self.optim.step()
        return {}

    def backward(self):
        """``local_iterations`` micro-batches, gradients *summed* (no 1/k)."""
        out = {}
        self.model.train()
        self.optim.zero_grad()
        its = []
        for _ in range(self.cache.get('local_iterations', 1)):
            batch, flags = self.trainer.data_handle.next_iter()
            it = self.trainer.iteration(batch)
            it['loss'].backward()
            its.append(it)
            out.update(**flags)
        return self.trainer.reduce_iteration(its), out

    def _ship(self, file_name, arrays):
        _tu.save_arrays(self.state['transferDirectory'] + _sep + file_name, _tu.as_object_array(arrays))

    def to_reduce(self):
        it, out = self.backward()
        out['grads_file'] = _conf.grads_file
        self._ship(out['grads_file'], _tu.extract_grads(self.model, dtype=self.dtype))
        out['reduce'] = True
        return it, out
