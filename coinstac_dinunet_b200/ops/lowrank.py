"""Device-side low-rank gradient compression (``csrc/lowrank.cu``): PowerSGD P/Q stages over every matrix of a model
in one launch each, and the rankDAD numerical core (Gram -> coefficient-space power iteration -> skinny GEMM ->
reconstruct-into-grad) with no cuSOLVER call, no host sync and no ``torch.matmul``.

Every function has a PyTorch oracle next to it (``*_reference``) that the GPU tests compare against and that runs
on CPU sites.
"""
import numpy as _np
import torch as _torch

from . import native as _nat


def _bump(n=1):
    from . import _count_launch
    _count_launch(n)


def _sp(device):
    return _nat.stream_ptr(device)


# ============================================================================================= PowerSGD
class PowerSGDPlan:
    """Descriptor tables for all compressible matrices (``ndim >= 2`` parameters viewed as ``[shape[0], -1]``) of a model
    whose gradients live in one flat arena.  ``p_numel`` / ``q_numel`` are the sizes of the factor buffers."""

    def __init__(self, params, offsets, rank, device):
        lib = _nat.lib()
        self.rank = int(rank)
        self.device = _torch.device(device)
        rows_per_tile, row_chunk = lib.coinn_psgd_rows_per_tile(), lib.coinn_psgd_row_chunk()
        self.mats, self.low = [], []          # (param index, n, m, g_off, p_off, q_off) / (param index, g_off, numel)
        p_off = q_off = 0
        for i, (p, off) in enumerate(zip(params, offsets)):
            n, m = (p.shape[0], p.numel() // p.shape[0]) if p.dim() >= 2 else (1, p.numel())
            if p.dim() >= 2 and min(n, m) > self.rank:      # a rank-r factorisation of a matrix with <= r rows saves nothing
                self.mats.append((i, n, m, off, p_off, q_off))
                p_off += n * self.rank
                q_off += m * self.rank
            else:
                self.low.append((i, off, p.numel()))
        self.p_numel, self.q_numel = p_off, q_off
        self.low_numel = sum(n for _, _, n in self.low)
        desc = _np.zeros(len(self.mats), dtype=[('g', '<i8'), ('p', '<i8'), ('q', '<i8'), ('n', '<i4'), ('m', '<i4')])
        assert desc.itemsize == lib.coinn_psgd_desc_size()
        row_tiles, col_tiles = [], []
        for k, (_, n, m, g, po, qo) in enumerate(self.mats):
            desc[k] = (g, po, qo, n, m)
            row_tiles += [(k, r0) for r0 in range(0, n, rows_per_tile)]
            col_tiles += [(k, c0, r0) for r0 in range(0, n, row_chunk) for c0 in range(0, m, 256)]
        to_dev = lambda a: _torch.from_numpy(a.view(_np.uint8).reshape(-1).copy()).to(self.device)
        self.desc = to_dev(desc)
        self.row_tiles = to_dev(_np.asarray(row_tiles, dtype=_np.int32).reshape(-1, 2))
        self.col_tiles = to_dev(_np.asarray(col_tiles, dtype=_np.int32).reshape(-1, 3))
        self.n_row_tiles, self.n_col_tiles = len(row_tiles), len(col_tiles)
        self.multi_chunk = any(n > row_chunk for _, n, *_ in self.mats)
        # gather / scatter tables for the rank-1 (bias, norm) gradients: arena <-> tail of the Q exchange buffer
        seg_g = _np.zeros(len(self.low), dtype=[('src', '<i8'), ('dst', '<i8'), ('len', '<i8')])
        seg_s = seg_g.copy()
        dst = self.q_numel
        for k, (_, off, numel) in enumerate(self.low):
            seg_g[k] = (off, dst, numel)
            seg_s[k] = (dst, off, numel)
            dst += numel
        self.seg_gather, self.seg_scatter = to_dev(seg_g), to_dev(seg_s)
        self.max_low = max([n for _, _, n in self.low] or [0])

    # ---- stages (all operate on flat fp32 buffers) --------------------------------------------------------------
    def orthogonalize(self, buf, which, eps=1e-8):
        _nat.check(_nat.lib().coinn_orthogonalize_batched(self.desc.data_ptr(), len(self.mats), buf.data_ptr(), int(which),
                                                          self.rank, float(eps), _sp(self.device)), 'orthogonalize_batched')
        _bump()

    def mq(self, G, E, Q, P, use_error):
        _nat.check(_nat.lib().coinn_psgd_mq(self.desc.data_ptr(), self.row_tiles.data_ptr(), self.n_row_tiles, G.data_ptr(),
                                            E.data_ptr(), Q.data_ptr(), P.data_ptr(), self.rank, int(use_error), _sp(self.device)),
                   'psgd_mq')
        _bump()

    def mtp(self, M, P, Q):
        if self.multi_chunk:
            Q[:self.q_numel].zero_()
        _nat.check(_nat.lib().coinn_psgd_mtp(self.desc.data_ptr(), self.col_tiles.data_ptr(), self.n_col_tiles, M.data_ptr(),
                                             P.data_ptr(), Q.data_ptr(), self.rank, _sp(self.device)), 'psgd_mtp')
        _bump()

    def reconstruct(self, G, E, P, Q, use_error):
        _nat.check(_nat.lib().coinn_psgd_reconstruct(self.desc.data_ptr(), self.row_tiles.data_ptr(), self.n_row_tiles,
                                                     G.data_ptr(), E.data_ptr(), P.data_ptr(), Q.data_ptr(), self.rank,
                                                     int(use_error), _sp(self.device)), 'psgd_reconstruct')
        _bump()

    def gather_low(self, G, exchange):
        if self.low:
            _nat.check(_nat.lib().coinn_segcopy(self.seg_gather.data_ptr(), len(self.low), self.max_low, G.data_ptr(),
                                                exchange.data_ptr(), 1.0, _sp(self.device)), 'segcopy[gather]')
            _bump()

    def scatter_low(self, exchange, G):
        if self.low:
            _nat.check(_nat.lib().coinn_segcopy(self.seg_scatter.data_ptr(), len(self.low), self.max_low, exchange.data_ptr(),
                                                G.data_ptr(), 1.0, _sp(self.device)), 'segcopy[scatter]')
            _bump()


def powersgd_round_reference(grads, errors, qs, rank, world_mean, use_error=True):
    """PyTorch oracle of one compressed round for a list of matrices (single site view): returns (approx, new errors,
    new qs).  ``world_mean(list_of_tensors)`` averages over sites in place (identity for one site)."""
    from ..distrib.powersgd import _orthogonalize
    Ms, Ps = [], []
    for g, e, q in zip(grads, errors, qs):
        M = g + e if use_error else g.clone()
        _orthogonalize(q)
        Ms.append(M)
        Ps.append(M @ q)
    world_mean(Ps)
    Qs = []
    for M, P in zip(Ms, Ps):
        _orthogonalize(P)
        Qs.append(M.t() @ P)
    world_mean(Qs)
    approx = [P @ Q.t() for P, Q in zip(Ps, Qs)]
    new_err = [M - a if use_error else _torch.zeros_like(M) for M, a in zip(Ms, approx)]
    return approx, new_err, Qs


# ============================================================================================== rankDAD
def lowrank_factor(B, C, rank, iters, tol, b_seg=None, c_seg=None, scale=1.0, out_left=None, out_right=None):
    """Top-``rank`` triplets of ``B @ C.T`` (B: [rowsB, n], C: [rowsC, n]) -> (left * sigma [rowsB, k], right [rowsC, k]).

    ``b_seg`` / ``c_seg`` = (base tensor, rows, n, kseg, seg_stride) describe column-block segmented operands (the
    all-gathered factors of S sites) instead of dense ``B`` / ``C``.  Five launches, no host sync:
    gram(B), gram(C), lowrank_eig, skinny_gemm(B, X), skinny_gemm(C, Y)."""
    lib = _nat.lib()
    if b_seg is None:
        B = B.contiguous().float()
        b_seg = (B, B.shape[0], B.shape[1], B.shape[1], 0)
    if c_seg is None:
        C = C.contiguous().float()
        c_seg = (C, C.shape[0], C.shape[1], C.shape[1], 0)
    (bt, rows_b, n, bk, bs), (ct, rows_c, n2, ck, cs) = b_seg, c_seg
    assert n == n2, (n, n2)
    dev = bt.device
    k = max(1, min(int(rank), rows_b, rows_c, n))
    grams = _torch.empty(2, n, n, dtype=_torch.float32, device=dev)
    coef = _torch.empty(2, n, k, dtype=_torch.float32, device=dev)
    chunk = lib.coinn_gram_rows_per_chunk()
    scratch = _torch.empty((-(-max(rows_b, rows_c) // chunk)) * n * n, dtype=_torch.float32, device=dev)
    sp = _sp(dev)
    # deterministic two-stage Gram sums (no atomics): every site repeats the re-compression and must get the same bits
    _nat.check(lib.coinn_gram_seg(bt.data_ptr(), int(bs), rows_b, n, bk, grams[0].data_ptr(), scratch.data_ptr(), sp), 'gram_seg[B]')
    _nat.check(lib.coinn_gram_seg(ct.data_ptr(), int(cs), rows_c, n, ck, grams[1].data_ptr(), scratch.data_ptr(), sp), 'gram_seg[C]')
    _nat.check(lib.coinn_lowrank_eig(grams[0].data_ptr(), grams[1].data_ptr(), n, k, int(iters), float(tol),
                                     coef[0].data_ptr(), coef[1].data_ptr(), sp), 'lowrank_eig')
    left = out_left if out_left is not None else _torch.empty(rows_b, k, dtype=_torch.float32, device=dev)
    right = out_right if out_right is not None else _torch.empty(rows_c, k, dtype=_torch.float32, device=dev)
    assert left.shape == (rows_b, k) and right.shape == (rows_c, k) and left.is_contiguous() and right.is_contiguous()
    _nat.check(lib.coinn_skinny_gemm_seg(bt.data_ptr(), int(bs), rows_b, n, bk, coef[0].data_ptr(), k, left.data_ptr(),
                                         float(scale), sp), 'skinny_gemm[B]')
    _nat.check(lib.coinn_skinny_gemm_seg(ct.data_ptr(), int(cs), rows_c, n, ck, coef[1].data_ptr(), k, right.data_ptr(),
                                         1.0, sp), 'skinny_gemm[C]')
    _bump(7)
    return left, right


def lowrank_factor_reference(B, C, rank, iters, tol):
    """PyTorch oracle of ``lowrank_factor`` - the same algorithm, step for step (runs anywhere, documents the kernel):
    with ``Gb = B^T B`` and ``Gc = C^T C`` (both ``[n, n]``) every left singular vector of ``G = B C^T`` is ``u = B x`` and
    ``G G^T u = B (Gc Gb x)``, so power iteration, normalisation (``x^T Gb x = 1``) and deflation against the previous vectors
    (``x -= (x_p^T Gb x) x_p``) all happen on n-vectors; ``sigma^2 = (Gb x)^T Gc (Gb x)``, the right vector is ``C (Gb x) / sigma``.
    Returns (left * sigma [rowsB, k], right [rowsC, k])."""
    B, C = B.double(), C.double()
    n = B.shape[1]
    k = max(1, min(int(rank), B.shape[0], C.shape[0], n))
    Gb, Gc = B.t() @ B, C.t() @ C
    X, Y, prev, prevw, sigma0 = [], [], [], [], None
    idx = _torch.arange(1, n + 1, dtype=_torch.float64)
    for c in range(k):
        x = 0.5 + 0.5 * _torch.sin(12.9898 * idx + 78.233 * (c + 1))
        for it in range(int(iters) + 1):
            for xp, wp in zip(prev, prevw):
                x = x - (wp @ x) * xp
            w = Gb @ x
            nrm2 = float(x @ w)
            inv = nrm2 ** -0.5 if nrm2 > 1e-30 else 0.0
            x, w = x * inv, w * inv
            z = Gc @ w
            sig2 = float(w @ z)
            if it == iters:
                break
            x = z
        sig = max(sig2, 0.0) ** 0.5
        sigma0 = sig if c == 0 else sigma0
        keep = sig > tol * max(sigma0, 1e-30) and sig > 0
        prev.append(x); prevw.append(w)
        X.append(x * sig if keep else _torch.zeros_like(x))
        Y.append(w / sig if keep else _torch.zeros_like(w))
    X, Y = _torch.stack(X, 1), _torch.stack(Y, 1)
    return (B @ X).float(), (C @ Y).float()


def dad_reconstruct(delta, act, weight_grad, bias_grad=None, scale=1.0):
    """``weight_grad[out, in] = scale * delta @ act[:in].T`` and, when ``act`` has one extra row (the bias column of the
    augmented activations), ``bias_grad[out] = scale * delta @ act[in]`` - written straight into the gradient arena."""
    out_f, in_f = weight_grad.shape
    k = delta.shape[1]
    assert delta.shape[0] == out_f and act.shape[1] == k and act.shape[0] in (in_f, in_f + 1)
    assert weight_grad.is_contiguous() and weight_grad.dtype == _torch.float32
    delta, act = delta.contiguous().float(), act.contiguous().float()
    bg = bias_grad if (bias_grad is not None and act.shape[0] == in_f + 1) else None
    _nat.check(_nat.lib().coinn_dad_reconstruct(delta.data_ptr(), act.data_ptr(), out_f, in_f, act.shape[0], k,
                                                weight_grad.data_ptr(), bg.data_ptr() if bg is not None else None,
                                                float(scale), _sp(delta.device)), 'dad_reconstruct')
    _bump()
