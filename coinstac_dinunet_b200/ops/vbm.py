"""Fused blocks of the VBM 3-D CNN on the hand-written kernels (``csrc/conv1_fused.cu``, ``csrc/conv3d_halo.cu``,
``csrc/conv3d_wgrad_*.cu``, ``csrc/conv3d_tma.cu``, ``csrc/vbm_fused.cu``).  Activations are channels-last
``[N, D, H, W, C]`` bf16.

One ``ConvBnReluPool`` block = Conv3d(k3, p1, no bias) -> BatchNorm3d (batch statistics) -> ReLU -> MaxPool3d(2).

Blocks 2-5 (stored conv output y):

    forward   conv (writes y; the halo kernel also emits the BatchNorm sums)  ->  bn+relu+pool (reads y, writes y/8)
    backward  pooled sums (reads p, dp -> dgamma, dbeta)  ->  apply (reads y, dp, writes dy)  ->  dgrad + wgrad
              ->  grad finalize (raw dW, dgamma, dbeta folded into .grad, accumulators re-zeroed)

Block 1 (C_in = 1, 543 MB of y at the benchmark shape) stores nothing at full resolution: the banded-Toeplitz tcgen05
convolution is recomputed in a statistics pass, a BN+ReLU+pool pass (pooled output + one arg-max code byte per value)
and the backward pass, where the gradient tile goes from registers to shared memory and straight into the weight-gradient
MMAs (``conv1_fused_*`` below).  The stored-y first-block kernels of round 1 (CUDA-core conv1, conv1_tc, stored-y Toeplitz)
were superseded by this path and removed.

For comparison the PyTorch chain is conv, BN-stat, BN-apply, ReLU, pool - each a full read+write - and its
channels-last-3d BatchNorm backward alone takes 40 ms per step on a B200 (profiles/r1_launches_torchmodules.txt).
"""
import ctypes as _C
import os as _os

import torch as _torch

from . import native as _nat

BF16 = _torch.bfloat16


def _bump(n=1):
    from . import _count_launch
    _count_launch(n)


def _sp(t):
    return _nat.stream_ptr(t.device)


def _chk(code, what):
    _nat.check(code, what)


# ------------------------------------------------------------------------------------ kernels
_SCRATCH = {}


def _scratch(param, name, numel):
    """Persistent zero-initialised fp32 accumulator tied to a parameter.  Its consumer kernel re-zeroes it in the launch
    that reads it (bn_finalize2 / conv_block_grad_finalize), so a training step issues no memset for it."""
    key = (param.data_ptr(), name, numel, param.device)
    buf = _SCRATCH.get(key)
    if buf is None:
        buf = _torch.zeros(numel, dtype=_torch.float32, device=param.device)
        _SCRATCH[key] = buf
    return buf


def _direct_ok(*params):
    from .linear import direct_grad_ok
    return all(direct_grad_ok(q) for q in params)


def bn_finalize2(stats, count, eps, momentum, running_mean, running_var, nbt=None):
    """bn_finalize + ``num_batches_tracked += 1`` + re-zeroing of the persistent ``stats`` accumulator, one launch."""
    C = stats.numel() // 2
    mean = _torch.empty(C, dtype=_torch.float32, device=stats.device)
    invstd = _torch.empty_like(mean)
    _chk(_nat.lib().coinn_bn_finalize2(stats.data_ptr(), mean.data_ptr(), invstd.data_ptr(),
                                       running_mean.data_ptr() if running_mean is not None else None,
                                       running_var.data_ptr() if running_var is not None else None,
                                       float(count), float(eps), float(momentum), C,
                                       nbt.data_ptr() if nbt is not None else None, 1, _sp(stats)), 'coinn_bn_finalize2')
    _bump()
    return mean, invstd


def conv_block_grad_finalize(dwt, acc, conv_w, gamma, beta, cin, cout, transposed):
    _chk(_nat.lib().coinn_conv_block_grad_finalize(dwt.data_ptr(), acc.data_ptr(), conv_w.grad.data_ptr(), gamma.grad.data_ptr(),
                                                   beta.grad.data_ptr(), cin, cout, int(transposed), _sp(dwt)),
         'coinn_conv_block_grad_finalize')
    _bump()
    from .linear import notify_grad_written
    notify_grad_written(conv_w, gamma, beta)         # direct mode: autograd never sees these gradients


def conv1_pad_input_hd(x):
    """[N,D,H,W] fp32/bf16 -> zero-padded bf16 row matrix [N*(H+2)*(D+2), Wq], d' fastest (conv1_fused.cu)."""
    N, D, H, W = x.shape
    x = x.contiguous() if x.dtype in (BF16, _torch.float32) else x.float().contiguous()
    rows, cols = _C.c_longlong(0), _C.c_int(0)
    _nat.lib().coinn_conv1_padded_shape(N, D, H, W, _C.byref(rows), _C.byref(cols))
    xp = _torch.empty((rows.value, cols.value), dtype=BF16, device=x.device)
    _chk(_nat.lib().coinn_conv1_pad_input_hd(x.data_ptr(), 1 if x.dtype == BF16 else 0, xp.data_ptr(), N, D, H, W, _sp(x)),
         'coinn_conv1_pad_input_hd')
    _bump()
    return xp


def _w27(weight):
    return weight.detach().float().reshape(16, 27).contiguous()


def conv1_fused_stats(xp, weight, shape, out=None):
    """-> stats[32]: per-channel sum and sum of squares of conv1(x) over the whole batch (nothing else is written)."""
    N, D, H, W = shape
    stats = out if out is not None else _torch.zeros(32, dtype=_torch.float32, device=xp.device)
    _chk(_nat.lib().coinn_conv1_fused_stats(xp.data_ptr(), _w27(weight).data_ptr(), stats.data_ptr(), N, D, H, W, _sp(xp)),
         'coinn_conv1_fused_stats')
    _bump()
    return stats


def conv1_fused_pool(xp, weight, mean, invstd, gamma, beta, shape):
    """-> (p [N,D/2,H/2,W/2,16] bf16 = maxpool2(relu(bn(conv1(x)))), code uint8 like p: arg-max position | 8*active)."""
    N, D, H, W = shape
    p = _torch.empty((N, D // 2, H // 2, W // 2, 16), dtype=BF16, device=xp.device)
    code = _torch.empty(p.shape, dtype=_torch.uint8, device=xp.device)
    _chk(_nat.lib().coinn_conv1_fused_pool(xp.data_ptr(), _w27(weight).data_ptr(), mean.data_ptr(), invstd.data_ptr(),
                                           gamma.data_ptr(), beta.data_ptr(), p.data_ptr(), code.data_ptr(), N, D, H, W, _sp(xp)),
         'coinn_conv1_fused_pool')
    _bump()
    return p, code


def conv1_fused_bwd(xp, weight, mean, invstd, gamma, beta, p, code, dp, shape, acc=None, dw=None):
    """-> (dW1 [16,1,3,3,3] fp32, dgamma, dbeta): BN/ReLU/pool backward + weight gradient with the conv output
    recomputed on the tensor cores and its gradient kept in shared memory (never in HBM)."""
    N, D, H, W = shape
    dp = dp.contiguous()
    acc = acc if acc is not None else _torch.zeros(32, dtype=_torch.float32, device=xp.device)
    _chk(_nat.lib().coinn_bn_pool_bwd_stats_pooled(p.data_ptr(), dp.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                                                   acc.data_ptr(), p.numel() // 16, 16, _sp(xp)), 'bn_pool_bwd_stats_pooled')
    dw = dw if dw is not None else _torch.zeros(16 * 27, dtype=_torch.float32, device=xp.device)
    _chk(_nat.lib().coinn_conv1_fused_bwd(xp.data_ptr(), _w27(weight).data_ptr(), mean.data_ptr(), invstd.data_ptr(),
                                          gamma.data_ptr(), acc.data_ptr(), dp.data_ptr(), code.data_ptr(), dw.data_ptr(),
                                          N, D, H, W, _sp(xp)), 'coinn_conv1_fused_bwd')
    _bump(2)
    return dw.view(16, 1, 3, 3, 3), acc[16:], acc[:16]


def bn_stats(y, out=None):
    """[..., C] bf16 -> stats[2C] = (sum, sumsq) over all leading dims."""
    C = y.shape[-1]
    stats = out if out is not None else _torch.zeros(2 * C, dtype=_torch.float32, device=y.device)
    _chk(_nat.lib().coinn_bn_stats(y.data_ptr(), stats.data_ptr(), y.numel() // C, C, _sp(y)), 'coinn_bn_stats')
    _bump()
    return stats


def bn_finalize(stats, count, eps, momentum, running_mean=None, running_var=None):
    C = stats.numel() // 2
    mean = _torch.empty(C, dtype=_torch.float32, device=stats.device)
    invstd = _torch.empty_like(mean)
    _chk(_nat.lib().coinn_bn_finalize(stats.data_ptr(), mean.data_ptr(), invstd.data_ptr(),
                                      running_mean.data_ptr() if running_mean is not None else None,
                                      running_var.data_ptr() if running_var is not None else None,
                                      float(count), float(eps), float(momentum), C, _sp(stats)), 'coinn_bn_finalize')
    _bump()
    return mean, invstd


def bn_relu_pool_fwd(y, mean, invstd, gamma, beta):
    N, D, H, W, C = y.shape
    p = _torch.empty((N, D // 2, H // 2, W // 2, C), dtype=BF16, device=y.device)
    _chk(_nat.lib().coinn_bn_relu_pool_fwd(y.data_ptr(), mean.data_ptr(), invstd.data_ptr(), gamma.data_ptr(),
                                           beta.data_ptr(), p.data_ptr(), N, D, H, W, C, _sp(y)), 'bn_relu_pool_fwd')
    _bump()
    return p


def bn_relu_pool_bwd(y, dp, mean, invstd, gamma, beta, p=None, acc=None):
    """-> (dy [N,D,H,W,C] bf16, dgamma [C], dbeta [C]).

    With the pooled forward output ``p`` given, pass A (dgamma / dbeta) runs on the pooled tensors only
    (xhat at the arg-max is recovered as (p - beta) / gamma): 1/8 of the bytes of the y-based pass."""
    N, D, H, W, C = y.shape
    acc = acc if acc is not None else _torch.zeros(2 * C, dtype=_torch.float32, device=y.device)
    dy = _torch.empty_like(y)
    args = (y.data_ptr(), dp.data_ptr(), mean.data_ptr(), invstd.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
            acc.data_ptr())
    if p is not None and _os.environ.get('COINN_BN_STATS_FROM_Y', '0') != '1':
        _chk(_nat.lib().coinn_bn_pool_bwd_stats_pooled(p.data_ptr(), dp.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                                                       acc.data_ptr(), p.numel() // C, C, _sp(y)), 'bn_pool_bwd_stats_pooled')
    else:
        _chk(_nat.lib().coinn_bn_relu_pool_bwd(*args, None, N, D, H, W, C, 0, _sp(y)), 'bn_relu_pool_bwd[A]')
    _chk(_nat.lib().coinn_bn_relu_pool_bwd(*args, dy.data_ptr(), N, D, H, W, C, 1, _sp(y)), 'bn_relu_pool_bwd[B]')
    _bump(2)
    return dy, acc[C:], acc[:C]


# ------------------------------------------------------------------------------------- convs
def _as_ncdhw(x_ndhwc):
    return x_ndhwc.permute(0, 4, 1, 2, 3)          # logical NCDHW, channels_last_3d strides (no copy)


def conv3d_fwd(x, weight, backend='auto', want_stats=False, stats_out=None):
    """x: [N,D,H,W,Cin] bf16, weight: [Cout,Cin,3,3,3] (any float dtype) -> y [N,D,H,W,Cout] bf16.
    ``want_stats``: -> (y, stats or None); stats = BatchNorm sums produced by the conv epilogue when the kernel can."""
    if want_stats:
        if backend in ('auto', 'tcgen05', 'fp8'):
            try:
                from .conv3d import conv3d_igemm_fwd
                return conv3d_igemm_fwd(x, weight, want_stats=True, stats_out=stats_out, fp8=(backend == 'fp8'))
            except ImportError:
                if backend == 'tcgen05':
                    raise
        return conv3d_fwd(x, weight, backend), None
    if backend in ('auto', 'tcgen05', 'fp8'):
        try:
            from .conv3d import conv3d_igemm_fwd
            return conv3d_igemm_fwd(x, weight, fp8=(backend == 'fp8'))
        except ImportError:
            if backend == 'tcgen05':
                raise
    w = weight.detach().to(BF16).contiguous(memory_format=_torch.channels_last_3d)
    y = _torch.nn.functional.conv3d(_as_ncdhw(x), w, padding=1)
    return y.permute(0, 2, 3, 4, 1).contiguous()


def conv3d_bwd(dy, x, weight, need_dx=True, backend='auto'):
    """-> (dx [N,D,H,W,Cin] bf16 or None, dW [Cout,Cin,3,3,3] fp32)"""
    if backend in ('auto', 'tcgen05', 'fp8'):
        try:
            from .conv3d import conv3d_igemm_bwd
            return conv3d_igemm_bwd(dy, x, weight, need_dx, fp8=(backend == 'fp8'))
        except ImportError:
            if backend == 'tcgen05':
                raise
    w = weight.detach().to(BF16).contiguous(memory_format=_torch.channels_last_3d)
    dx, dw, _ = _torch.ops.aten.convolution_backward(
        _as_ncdhw(dy), _as_ncdhw(x), w, None, [1, 1, 1], [1, 1, 1], [1, 1, 1], False, [0, 0, 0], 1,
        [bool(need_dx), True, False])
    if dx is not None:
        dx = dx.permute(0, 2, 3, 4, 1).contiguous()
    return dx, dw.float()


# --------------------------------------------------------------------------------- autograd
class ConvBnReluPoolFn(_torch.autograd.Function):
    """One fused VBM block.  ``x``: [N,D,H,W,Cin] bf16 (or [N,D,H,W] fp32/bf16 for the first block).

    When the three parameters already own fp32 ``.grad`` buffers (the DistArena gradient arena) the block runs in
    "direct" mode: BatchNorm sums, BN-backward sums and the raw weight gradient live in persistent accumulators that their
    consumer kernels re-zero, ``num_batches_tracked`` is bumped by ``bn_finalize2`` and one ``conv_block_grad_finalize``
    launch folds everything into the ``.grad`` buffers - the block returns ``None`` for its parameter gradients and a
    training step issues no memset / permute / AccumulateGrad launches for it."""

    @staticmethod
    def forward(ctx, x, conv_w, gamma, beta, running_mean, running_var, eps, momentum, training, backend, nbt=None):
        first = x.dim() == 4
        g, b = gamma.detach().float().contiguous(), beta.detach().float().contiguous()
        cout = conv_w.shape[0]
        direct = bool(training) and x.is_cuda and _direct_ok(conv_w, gamma, beta) and \
            (backend in ('auto', 'tcgen05', 'fp8') or first)
        ctx.direct, ctx.params = direct, (conv_w, gamma, beta)
        stats_buf = _scratch(conv_w, 'stats', 2 * cout) if direct else None

        def finalize(stats, count):
            if direct:
                return bn_finalize2(stats, count, eps, momentum, running_mean, running_var, nbt)
            if nbt is not None:
                nbt.add_(1)
            return bn_finalize(stats, count, eps, momentum, running_mean, running_var)

        if first:
            if conv_w.shape[0] != 16 or conv_w.shape[1] != 1:
                raise ValueError('the fused first block is instantiated for Conv3d(1 -> 16); other first layers go through '
                                 'the generic blocks (ops.nativize pads C_in)')
            shape = tuple(x.shape)
            xp = conv1_pad_input_hd(x)
            if training:
                mean, invstd = finalize(conv1_fused_stats(xp, conv_w, shape, out=stats_buf), x.numel())
            else:
                mean, invstd = running_mean.float(), (running_var.float() + eps).rsqrt()
            p, code = conv1_fused_pool(xp, conv_w, mean, invstd, g, b, shape)
            ctx.save_for_backward(xp, conv_w, code, mean, invstd, g, b, p)
            ctx.first, ctx.backend, ctx.training, ctx.fused_shape = first, backend, training, shape
            return p
        ctx.fused_shape = None
        res = conv3d_fwd(x, conv_w, backend, want_stats=bool(training), stats_out=stats_buf)
        y, stats = res if training else (res, None)
        N, D, H, W, C = y.shape
        if training:
            if stats is None:
                stats = bn_stats(y, out=stats_buf)
            mean, invstd = finalize(stats, N * D * H * W)
        else:
            mean = running_mean.float()
            invstd = (running_var.float() + eps).rsqrt()
        p = bn_relu_pool_fwd(y, mean, invstd, g, b)
        ctx.save_for_backward(x, conv_w, y, mean, invstd, g, b, p)
        ctx.first, ctx.backend, ctx.training = first, backend, training
        return p

    @staticmethod
    def backward(ctx, dp):
        if not ctx.training:   # eval-mode BN has no batch-statistics terms; not a training path
            raise RuntimeError('ConvBnReluPoolFn.backward is only defined for training-mode BatchNorm')
        conv_w_p, gamma_p, beta_p = ctx.params
        cout, cin = conv_w_p.shape[0], conv_w_p.shape[1]
        direct = ctx.direct and _direct_ok(conv_w_p, gamma_p, beta_p)
        if direct and not ctx.first:
            from .conv3d_wgrad import _SUPPORTED
            direct = (cin, cout) in _SUPPORTED
        acc = _scratch(conv_w_p, 'acc', 2 * cout) if direct else None
        none = (None,) * 7
        if ctx.fused_shape is not None:
            xp, conv_w, code, mean, invstd, g, b, p = ctx.saved_tensors
            dwbuf = _scratch(conv_w_p, 'dw', 27 * cin * cout) if direct else None
            dw, dgamma, dbeta = conv1_fused_bwd(xp, conv_w, mean, invstd, g, b, p, code, dp, ctx.fused_shape, acc=acc, dw=dwbuf)
            if direct:
                conv_block_grad_finalize(dwbuf, acc, conv_w_p, gamma_p, beta_p, cin, cout, transposed=False)
                return (None, None, None, None) + none
            return (None, dw.to(conv_w.dtype), dgamma.to(g.dtype), dbeta.to(b.dtype)) + none
        x, conv_w, y, mean, invstd, g, b, p = ctx.saved_tensors
        dy, dgamma, dbeta = bn_relu_pool_bwd(y, dp.contiguous(), mean, invstd, g, b, p=p, acc=acc)
        if direct:
            from .conv3d import conv3d_igemm_bwd
            dwbuf = _scratch(conv_w_p, 'dw', 27 * cin * cout)
            dx, _ = conv3d_igemm_bwd(dy, x, conv_w, need_dx=ctx.needs_input_grad[0], raw_dw=dwbuf.view(27 * cin, cout),
                                     fp8=(ctx.backend == 'fp8'))
            conv_block_grad_finalize(dwbuf, acc, conv_w_p, gamma_p, beta_p, cin, cout, transposed=True)
            return (dx, None, None, None) + none
        else:
            dx, dw = conv3d_bwd(dy, x, conv_w, need_dx=ctx.needs_input_grad[0], backend=ctx.backend)
        return (dx, dw.to(conv_w.dtype), dgamma.to(g.dtype), dbeta.to(b.dtype)) + none
