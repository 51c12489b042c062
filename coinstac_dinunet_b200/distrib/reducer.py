"""Aggregator-side half of dSGD: load every site's payload, average, ship it back.

File wire format and JSON keys as in coinstac_dinunet/distrib/reducer.py:11-54.  The mean is
computed on one flat ``[S, N]`` buffer (one H2D, one reduction, one D2H) in the payload
dtype - numerically the reference's per-parameter ``mean(0)`` without its
3·P tiny launches.  ``mp_pool`` may be ``None`` (sequential loads; quirk §8.5-13).
"""
import os as _os
from functools import partial as _partial

import numpy as _np
import torch as _torch

from .. import config as _conf
from ..utils import tensorutils as _tu


def _multi_load(file_key, state, site, site_vars):
    return _tu.load_arrays(state['baseDirectory'] + _os.sep + site + _os.sep + site_vars[file_key])


class COINNReducer:
    def __init__(self, trainer, mp_pool, **kw):
        self.trainer = trainer
        self.cache = trainer.cache
        self.input = trainer.input
        self.state = trainer.state
        self.pool = mp_pool
        self.dtype = f"float{self.cache.get('precision_bits', 32)}"
        self.device = trainer.device.get('gpu', _torch.device('cpu'))

    def _load(self, file_key):
        job = _partial(_multi_load, file_key, self.state)
        sites = list(self.input.items())
        if self.pool is not None:
            return list(self.pool.starmap(job, sites))
        return [job(site, site_vars) for site, site_vars in sites]

    def _average(self, file_key):
        per_site = [list(arrs) for arrs in self._load(file_key)]
        if not per_site or not per_site[0]:
            return []
        shapes = [_np.shape(a) for a in per_site[0]]
        sizes = [int(_np.prod(s)) for s in shapes]
        stacked = _np.stack([
            _np.concatenate([_np.asarray(a, dtype=self.dtype).reshape(-1) for a in arrs]) for arrs in per_site
        ])
        mean = _torch.from_numpy(stacked).to(self.device, non_blocking=True).mean(0)
        flat = mean.cpu().numpy().astype(self.dtype)
        bounds = _np.cumsum([0] + sizes)
        return [flat[bounds[i]:bounds[i + 1]].reshape(shapes[i]) for i in range(len(sizes))]

    def _ship(self, file_name, arrays):
        _tu.save_arrays(self.state['transferDirectory'] + _os.sep + file_name, _tu.as_object_array(arrays))

    def reduce(self):
        """Average every site's gradients and hand the mean to all sites."""
        out = {'avg_grads_file': _conf.avg_grads_file}
        self._ship(out['avg_grads_file'], self._average('grads_file'))
        out['update'] = True
        return out
