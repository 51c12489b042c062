"""PowerSGD: rank-r gradient compression with error feedback and warm start.

Protocol parity with coinstac_dinunet/distrib/powersgd/__init__.py:15-219 (file names, JSON
keys ``start_power_iter`` / ``powerSGD_phase`` / ``powerSGD_{P,Q}_file[_AGG]`` /
``rank1_grads_file`` / ``rank_1_grads_file_AGG``, two round trips per optimizer step after
``start_powerSGD_iter`` vanilla dSGD steps).

Differences (documented):
* any parameter with ``ndim >= 2`` is compressed as the matrix ``[shape[0], -1]`` (conv kernels
  work; the reference only handles 2-D weights);
* the error memory is ``M - P·Qᵀ`` with ``M = grad + previous error`` (textbook PowerSGD); the
  reference subtracts from the raw gradient;
* Gram-Schmidt runs through ``ops.orthogonalize`` (one CTA per matrix on sm_100a) when the
  native extension is loaded; the PyTorch path below is the oracle.
"""
import os as _os
from collections import OrderedDict as _Dict

import torch as _torch

from ... import config as _conf
from ...utils import tensorutils as _tu
from ..learner import COINNLearner as _COINNLearner
from ..reducer import COINNReducer as _COINNReducer

_sep = _os.sep


def _orthogonalize(matrix, epsilon=1e-8):
    """In-place modified Gram-Schmidt over the columns of a tall ``[m, r]`` matrix.

    ``epsilon`` guards the normalisation against vanishing columns.  Column ``i`` is
    normalised, then its component is removed from all later columns in one rank-1 update.
    """
    cols = matrix.shape[1]
    for i in range(cols):
        q = matrix[:, i:i + 1]
        q.div_(q.norm() + epsilon)
        if i + 1 < cols:
            tail = matrix[:, i + 1:]
            tail.sub_(q @ (q.t() @ tail))
    return matrix


def _native_orthogonalize(matrix, epsilon=1e-8):
    try:
        from ... import ops as _ops
        if matrix.is_cuda and _ops.native_available():
            return _ops.orthogonalize_(matrix, epsilon)
    except Exception:
        pass
    return _orthogonalize(matrix, epsilon)


class PowerSGDState:
    """Persistent (in ``cache``) compression state of one site."""

    def __init__(self):
        self.error_dict = _Dict()
        self.p_memory_dict = _Dict()
        self.q_memory_dict = _Dict()
        self.rank1_tensors = _Dict()
        self.high_rank_tensors = _Dict()
        self.iter = 0


def _as_matrix(t):
    return t.reshape(t.shape[0], -1)


class PowerSGDLearner(_COINNLearner):
    def __init__(self, **kw):
        super().__init__(**kw)
        c = self.cache
        self.matrix_approximation_rank = c.setdefault('matrix_approximation_rank', 1)
        self.start_powerSGD_iter = c.setdefault('start_powerSGD_iter', 10)
        self.use_error_feedback = c.setdefault('use_error_feedback', True)
        self.warm_start = c.setdefault('warm_start', True)
        self.seed = c.get('seed') or 0
        self.powerSGD_state = c.setdefault('powerSGD_state', PowerSGDState())

    @property
    def _warming_up(self):
        return self.powerSGD_state.iter < self.start_powerSGD_iter

    def _np(self, t):
        return t.detach().float().cpu().numpy().astype(self.dtype)

    # ------------------------------------------------------------------- apply
    def step(self) -> dict:
        st = self.powerSGD_state
        if self._warming_up:
            st.iter += 1
            return super().step()

        base = self.state['baseDirectory'] + _sep
        avg_qs = _tu.load_arrays(base + self.input['powerSGD_Q_file_AGG'])
        for key, q in zip(st.p_memory_dict, avg_qs):
            st.q_memory_dict[key] = _torch.from_numpy(q).float().to(self.device)
        low = [_torch.from_numpy(g).float().to(self.device)
               for g in _tu.load_arrays(base + self.input['rank_1_grads_file_AGG'])]
        low.reverse()  # pop() in parameter order

        for key, p in self.model.named_parameters():
            if p.grad is None:
                continue
            if p.grad.dim() <= 1:
                p.grad = low.pop().reshape(p.shape).to(p.dtype)
                continue
            approx = st.p_memory_dict[key] @ st.q_memory_dict[key].t()
            if self.use_error_feedback:
                st.error_dict[key] = st.high_rank_tensors[key] - approx
            p.grad = approx.reshape(p.shape).to(p.dtype)
        st.high_rank_tensors = _Dict()
        self.optim.step()
        st.iter += 1
        return {}

    # ----------------------------------------------------------------- compress
    def _prepare_parameters(self):
        """backward; split grads into 1-D (sent raw) and matrices (compressed); ``P = M·Q``."""
        it, out = self.backward()
        st = self.powerSGD_state
        for key, p in self.model.named_parameters():
            g = p.grad.detach().float().clone()
            if p.dim() <= 1:
                st.rank1_tensors[key] = g
            else:
                st.high_rank_tensors[key] = _as_matrix(g)

        r = self.matrix_approximation_rank
        for key, M in st.high_rank_tensors.items():
            if self.use_error_feedback:
                if key in st.error_dict:
                    M += st.error_dict[key]
                else:
                    st.error_dict[key] = _torch.zeros_like(M)
            if not self.warm_start or key not in st.p_memory_dict:
                # same seed on every site -> identical Q without communication
                gen = _torch.Generator(device='cpu').manual_seed(int(self.seed) + int(st.iter))
                st.q_memory_dict[key] = _torch.randn(M.shape[1], r, generator=gen, dtype=M.dtype).to(M.device)
            _native_orthogonalize(st.q_memory_dict[key])
            st.p_memory_dict[key] = M @ st.q_memory_dict[key]
        return it, out

    def to_reduce(self):
        if self._warming_up:
            it, out = super().to_reduce()
            out['start_power_iter'] = False
            return it, out

        st, it, out = self.powerSGD_state, {}, {}
        phase = self.input.get('powerSGD_phase', 'phase_P_sync')
        if phase == 'phase_P_sync':
            it, out = self._prepare_parameters()
            out['powerSGD_P_file'] = f"powerSGD_P_{_conf.grads_file}"
            self._ship(out['powerSGD_P_file'], [self._np(p) for p in st.p_memory_dict.values()])

        elif phase == 'phase_Q_sync':
            out['rank1_grads_file'] = f"rank1_{_conf.grads_file}"
            self._ship(out['rank1_grads_file'], [self._np(g) for g in st.rank1_tensors.values()])
            st.rank1_tensors = _Dict()

            avg_ps = _tu.load_arrays(self.state['baseDirectory'] + _sep + self.input['powerSGD_P_file_AGG'])
            for key, p in zip(list(st.p_memory_dict), avg_ps):
                st.p_memory_dict[key] = _native_orthogonalize(_torch.from_numpy(p).float().to(self.device))
            for key, M in st.high_rank_tensors.items():
                st.q_memory_dict[key] = M.t() @ st.p_memory_dict[key]
            out['powerSGD_Q_file'] = f"powerSGD_Q_{_conf.grads_file}"
            self._ship(out['powerSGD_Q_file'], [self._np(q) for q in st.q_memory_dict.values()])

        out['start_power_iter'] = True
        out['reduce'] = True
        return it, out


class PowerSGDReducer(_COINNReducer):
    def reduce(self):
        """Warm-up: plain dSGD. Afterwards alternate: mean(P) -> Q-sync; mean(Q, rank-1) -> update."""
        site = next(iter(self.input.values()))
        if not site['start_power_iter']:
            return super().reduce()

        out = {}
        if site.get('powerSGD_P_file'):
            out['powerSGD_P_file_AGG'] = f"powerSGD_P_{_conf.avg_grads_file}"
            self._ship(out['powerSGD_P_file_AGG'], self._average('powerSGD_P_file'))
            out['powerSGD_phase'] = 'phase_Q_sync'
        elif site.get('powerSGD_Q_file'):
            out['rank_1_grads_file_AGG'] = f"rank1_AGG_{_conf.avg_grads_file}"
            self._ship(out['rank_1_grads_file_AGG'], self._average('rank1_grads_file'))
            out['powerSGD_Q_file_AGG'] = f"powerSGD_Q_{_conf.avg_grads_file}"
            self._ship(out['powerSGD_Q_file_AGG'], self._average('powerSGD_Q_file'))
            out['powerSGD_phase'] = 'phase_P_sync'
            out['update'] = True
        return out
