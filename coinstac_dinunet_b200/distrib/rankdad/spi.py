"""rankDAD building blocks: structured power iteration and the hook-based module wrapper.

Parity: coinstac_dinunet/distrib/rankdad/spi.py:9-250 (``power_iteration_BC`` contract,
``DADParallel.{train,eval,dad_backward,synced_param_update}``, ``dad_data.npy`` payload of
``[delta[out,r,1], act[in,r]]`` pairs in reverse layer order).

B200-first redesign of the numerical core: the reference extracts singular triplets one at a
time with deflation and several host syncs per vector (``isnan``, ``== 0.0``; spi.py:63,80).
Here all ``rank`` vectors advance together (block / subspace iteration with a thin QR per
sweep, then one small SVD), there is no data-dependent control flow and therefore no
device->host sync; the whole per-layer compression is a handful of GEMMs that batch well.
"""
import os as _os

import numpy as _np
import torch as _torch

from ...utils import tensorutils as _tu

_SKIP_NORM_Layers = (_torch.nn.BatchNorm1d, _torch.nn.BatchNorm2d, _torch.nn.BatchNorm3d,
                     _torch.nn.LayerNorm, _torch.nn.GroupNorm)
_DAD_LAYERS = (_torch.nn.Linear,)


def power_iteration_BC(B, C, rank, numiterations, tol, generator=None):
    """Top-``rank`` singular triplets of ``B @ C.T`` without forming it.

    ``B`` is ``[rowsB, n]``, ``C`` is ``[rowsC, n]`` (``n`` = samples).  Returns
    ``(left·σ  [rowsB, k], right [rowsC, k])`` so that ``left_sigma @ right.T ≈ B @ C.T``.
    Components whose singular value falls below ``tol · σ_max`` are zeroed (the reference drops
    them by early exit); ``k = min(rank, rowsB, rowsC, n)`` is static, which keeps shapes
    graph-capturable.
    """
    rows_b, n = B.shape
    rows_c = C.shape[0]
    k = max(1, min(int(rank), rows_b, rows_c, n))
    Bf, Cf = B.float(), C.float()
    # G Gᵀ = B (CᵀC) Bᵀ ; keep whichever inner Gram matrix is smaller
    if rows_c > n:
        inner = Cf.t() @ Cf                       # [n, n]

        def ggt(X):
            return Bf @ (inner @ (Bf.t() @ X))
    else:
        bct = Bf @ Cf.t()                         # [rowsB, rowsC] (small side)

        def ggt(X):
            return bct @ (bct.t() @ X)

    if generator is None:
        generator = _torch.Generator(device='cpu').manual_seed(rows_b * 7919 + rows_c * 104729 + n)
    X = _torch.randn(rows_b, k, generator=generator, dtype=_torch.float32).to(B.device)
    X, _ = _torch.linalg.qr(X)
    for _ in range(max(int(numiterations), 1)):
        X, _ = _torch.linalg.qr(ggt(X))
    # Gᵀ X = C (Bᵀ X): [rowsC, k]; its thin SVD rotates X onto the singular directions
    W = Cf @ (Bf.t() @ X)
    U, S, Vh = _torch.linalg.svd(W, full_matrices=False)
    left = X @ Vh.t()
    keep = (S > tol * S.max().clamp_min(1e-30)).to(S.dtype)
    S = S * keep
    return left * S.unsqueeze(0), U * keep.unsqueeze(0)


def _dad_trainable_module(module):
    """Leaf with parameters that is not a normalisation layer."""
    if isinstance(module, _SKIP_NORM_Layers):
        return False
    return any(True for _ in module.parameters())


def _mm_flatten(*tensors):
    """Collapse all leading dims so every tensor is ``[samples, features]``."""
    return [t.reshape(-1, t.shape[-1]) if t.dim() > 2 else t for t in tensors]


class DADParallel(_torch.nn.Module):
    """Wraps a model; in train mode records, per DAD layer, the layer input ``A`` and the
    gradient w.r.t. the layer output ``Δ`` so the weight gradient ``Δᵀ·A`` can be exchanged in
    factored low-rank form.

    ``bias_augment`` (default on): ``A`` gets a trailing column of ones so the bias gradient
    ``Δᵀ·1`` is carried exactly by the same factors (the reference approximates it by summing
    the left factors, spi.py:206).
    """

    def __init__(self, module, cache=None, input=None, state=None, device=None, dtype='float32', **kw):
        super().__init__()
        self.module = module.module if isinstance(module, DADParallel) else module
        self.cache, self.input, self.state = cache, input, state
        self.device, self.dtype = device, dtype
        self._is_dad_module = self.cache.setdefault('is_dad_module', {})
        self.rank = self.cache.setdefault('dad_reduction_rank', 10)
        self.num_pow_iters = self.cache.setdefault('dad_num_pow_iters', 5)
        self.dad_tol = self.cache.setdefault('dad_tol', 1e-3)
        self.bias_augment = self.cache.setdefault('dad_bias_augment', True)
        self._handles = []
        self._reset()

    # ---------------------------------------------------------------- plumbing
    def _reset(self):
        self._activations, self._local_grads = {}, {}

    def _leaves(self, reverse=False):
        """(name, module) of every leaf in definition order (or reversed)."""
        found = [(n, m) for n, m in self.module.named_modules() if n and not any(True for _ in m.children())]
        return found[::-1] if reverse else found

    def dad_layers(self, reverse=False):
        return [(n, m) for n, m in self._leaves(reverse) if self._is_dad_module.get(n)]

    def plain_parameters(self):
        """Parameters *not* covered by rankDAD factors (norm layers, convs, ...).  They are
        averaged densely so no parameter is left unsynchronised (reference quirk §8.5-9)."""
        dad = {id(p) for _, m in self.dad_layers() for p in m.parameters()}
        return [p for p in self.module.parameters() if id(p) not in dad]

    def _hook(self):
        self._unhook()
        for name, m in self._leaves():
            is_dad = isinstance(m, _DAD_LAYERS) and _dad_trainable_module(m)
            self._is_dad_module[name] = is_dad
            if not (is_dad and self.training):
                continue

            def fwd(mod, args, output, key=name):
                if args and args[0] is not None:
                    self._activations[key] = args[0].detach()

            def bwd(mod, grad_input, grad_output, key=name):
                if grad_output and grad_output[0] is not None:
                    self._local_grads[key] = grad_output[0].detach()

            self._handles.append(m.register_forward_hook(fwd))
            self._handles.append(m.register_full_backward_hook(bwd))

    def _unhook(self):
        for h in self._handles:
            h.remove()
        self._handles = []

    def train(self, mode=True):
        super().train(mode)
        self.module.train(mode)
        if mode:
            self._hook()
        else:
            self._unhook()
        return self

    def eval(self):
        return self.train(False)

    def forward(self, *inputs, **kwargs):
        if self.training:
            self._reset()
        return self.module(*inputs, **kwargs)

    def state_dict(self, *a, **kw):  # checkpoints stay wrapper-agnostic
        return self.module.state_dict(*a, **kw)

    def load_state_dict(self, *a, **kw):
        return self.module.load_state_dict(*a, **kw)

    # ------------------------------------------------------------- site -> remote
    def _factors(self, name, module):
        delta, act = _mm_flatten(self._local_grads[name].float(), self._activations[name].float())
        if self.bias_augment and getattr(module, 'bias', None) is not None:
            act = _torch.cat([act, act.new_ones(act.shape[0], 1)], dim=1)
        return power_iteration_BC(delta.t(), act.t(), self.rank, self.num_pow_iters, self.dad_tol)

    def dad_backward(self):
        """Compress every DAD layer (last layer first) and write ``dad_data.npy`` (+ the dense
        gradients of the remaining parameters in ``dad_plain_grads.npy``)."""
        out, payload = {}, []
        for name, m in self.dad_layers(reverse=True):
            g, a = self._factors(name, m)
            payload.append([g.unsqueeze(-1).cpu().numpy().astype(self.dtype), a.cpu().numpy().astype(self.dtype)])
        out['dad_data'] = 'dad_data.npy'
        box = _np.empty(len(payload), dtype=object)
        for i, pair in enumerate(payload):
            inner = _np.empty(2, dtype=object)
            inner[0], inner[1] = pair
            box[i] = inner
        _tu.save_arrays(self.state['transferDirectory'] + _os.sep + out['dad_data'], box)

        plain = [(_torch.zeros_like(p) if p.grad is None else p.grad).detach().float().cpu().numpy()
                 .astype(self.dtype) for p in self.plain_parameters()]
        if plain:
            out['dad_plain_grads'] = 'dad_plain_grads.npy'
            _tu.save_arrays(self.state['transferDirectory'] + _os.sep + out['dad_plain_grads'],
                            _tu.as_object_array(plain))
        return out

    # ------------------------------------------------------------- remote -> site
    def synced_param_update(self):
        """Rebuild gradients from the reduced factors: ``W.grad = Δᵣ·Aᵣᵀ`` (``[out, in]``)."""
        base = self.state['baseDirectory'] + _os.sep
        pairs = list(_tu.load_arrays(base + self.input['reduced_dad_data']))
        for (name, m), pair in zip(self.dad_layers(reverse=True), pairs):
            delta = _torch.from_numpy(_np.asarray(pair[0], dtype=_np.float32)).to(self.device)
            delta = delta.reshape(delta.shape[0], -1)                    # [out, k]
            act = _torch.from_numpy(_np.asarray(pair[1], dtype=_np.float32)).to(self.device)  # [in(+1), k]
            full = delta @ act.t()                                       # [out, in(+1)]
            has_bias = getattr(m, 'bias', None) is not None
            if has_bias and act.shape[0] == m.weight.shape[1] + 1:
                m.weight.grad = full[:, :-1].contiguous().to(m.weight.dtype)
                m.bias.grad = full[:, -1].contiguous().to(m.bias.dtype)
            else:
                m.weight.grad = full.contiguous().to(m.weight.dtype)
                if has_bias:
                    m.bias.grad = delta.sum(1).to(m.bias.dtype)          # reference behaviour
        if self.input.get('reduced_dad_plain_grads'):
            dense = _tu.load_arrays(base + self.input['reduced_dad_plain_grads'])
            for p, g in zip(self.plain_parameters(), dense):
                p.grad = _torch.from_numpy(_np.asarray(g, dtype=_np.float32)).to(self.device).reshape(p.shape) \
                    .to(p.dtype)
