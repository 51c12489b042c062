"""MX-FP8 block-scaled compute path (``csrc/mxfp8.cu``): e4m3 elements with one ue8m0 (power-of-two) scale per 32
elements along the contraction axis, multiplied by ``tcgen05.mma.kind::mxf8f6f4.block_scale`` with the scale factors
staged in TMEM (BASELINE config 4).  fp32 master weights stay in the fused reduce+Adam arena;
only the GEMM operands are quantised.

``quantize_mx`` / ``dequantize_mx`` / ``gemm_mxfp8`` are the kernel-level API; ``Fp8LinearFn`` runs the three GEMMs of a
Linear layer (forward, dgrad, wgrad) in MX-FP8, re-quantising every operand along ITS contraction axis (the transposed
operands of dgrad / wgrad need their own scales - SURVEY §7.3 hard part 5).
"""
import torch as _torch

from . import native as _nat


def _bump(n=1):
    from . import _count_launch
    _count_launch(n)


def quantize_mx(x):
    """x: [R, K] fp32 / bf16 (CUDA) -> (q uint8 [R, Kp], sf int32 [R, Kp // 128]); Kp = K rounded up to 128 (zero padded).
    Byte j of word g of row r is the ue8m0 exponent of elements [128 g + 32 j, 128 g + 32 j + 32) of that row."""
    assert x.is_cuda and x.dim() == 2
    if x.dtype not in (_torch.float32, _torch.bfloat16):
        x = x.float()
    x = x.contiguous()
    R, K = x.shape
    Kp = (K + 127) // 128 * 128
    q = _torch.empty((R, Kp), dtype=_torch.uint8, device=x.device)
    sf = _torch.empty((R, Kp // 128), dtype=_torch.int32, device=x.device)
    _nat.check(_nat.lib().coinn_quantize_mx(x.data_ptr(), 1 if x.dtype == _torch.bfloat16 else 0, q.data_ptr(), sf.data_ptr(),
                                            R, K, Kp, _nat.stream_ptr(x.device)), 'coinn_quantize_mx')
    _bump()
    return q, sf


def dequantize_mx(q, sf, K=None):
    """PyTorch oracle: (q, sf) -> fp32 [R, K]."""
    R, Kp = q.shape
    vals = q.view(_torch.float8_e4m3fn).float().view(R, Kp // 32, 32)
    exps = sf.view(_torch.uint8).view(R, Kp // 32).float()
    out = (vals * _torch.exp2(exps - 127.0).unsqueeze(-1)).view(R, Kp)
    return out[:, :K] if K is not None else out


def quantize_mx_reference(x):
    """PyTorch oracle of ``quantize_mx`` (same scale rule: smallest power of two >= amax / 448, clamped to [2^-126, 2^127])."""
    x = x.float()
    R, K = x.shape
    Kp = (K + 127) // 128 * 128
    xp = _torch.nn.functional.pad(x, (0, Kp - K)).view(R, Kp // 32, 32)
    amax = xp.abs().amax(-1)
    e = _torch.ceil(_torch.log2(amax / 448.0)).clamp(-126, 127)
    e = _torch.where(amax > 0, e, _torch.zeros_like(e))
    q = (xp / _torch.exp2(e).unsqueeze(-1)).clamp(-448, 448).to(_torch.float8_e4m3fn)
    return q.view(R, Kp).view(_torch.uint8), (e + 127).to(_torch.uint8).view(R, Kp // 128, 4).contiguous().view(_torch.int32).view(R, Kp // 128)


def gemm_mxfp8(aq, asf, bq, bsf, bias=None, relu=False, out_dtype=_torch.float32, bias_mode=1, split_k=None, out=None):
    """``dequant(a)[M, K] @ dequant(b)[N, K].T`` -> [M, N] on the block-scaled tensor-core path."""
    M, K = aq.shape
    N, K2 = bq.shape
    assert K == K2 and K % 128 == 0 and aq.dtype == _torch.uint8 and bq.dtype == _torch.uint8
    assert asf.shape == (M, K // 128) and bsf.shape == (N, K // 128)
    if split_k is None:
        tiles = ((M + 127) // 128) * ((N + 127) // 128)
        split_k = max(1, min(K // 128, 148 // tiles)) if tiles * 2 <= 148 else 1
    if split_k > 1:
        c = _torch.zeros((M, N), dtype=_torch.float32, device=aq.device)
        code = 1
    else:
        c = out if out is not None else _torch.empty((M, N), dtype=out_dtype, device=aq.device)
        code = 0 if c.dtype == _torch.bfloat16 else 1
    b32 = bias.float().contiguous() if (bias is not None and split_k <= 1) else None
    _nat.check(_nat.lib().coinn_gemm_mxfp8_tn(aq.data_ptr(), asf.data_ptr(), bq.data_ptr(), bsf.data_ptr(), c.data_ptr(),
                                              b32.data_ptr() if b32 is not None else None, M, N, K, N, code,
                                              int(relu and split_k <= 1), int(bias_mode), int(split_k), _nat.stream_ptr(aq.device)),
               'coinn_gemm_mxfp8_tn')
    _bump()
    if split_k > 1:
        if bias is not None:
            c = c + (bias.float() if bias_mode == 1 else bias.float().unsqueeze(1))
        if relu:
            c = c.relu_()
        c = c if out_dtype == _torch.float32 else c.to(out_dtype)
        if out is not None:
            out.copy_(c)
            return out
    return c


class Fp8LinearFn(_torch.autograd.Function):
    """y = relu?(x @ W^T + b) with forward, dgrad and wgrad on the MX-FP8 block-scaled GEMM.
    forward  y  = Q_k(x)  . Q_k(W)^T            contraction over in_features
    dgrad    dx = Q_n(dy) . Q_n(W^T)^T          contraction over out_features (W re-quantised along its other axis)
    wgrad    dW = Q_m(dy^T) . Q_m(x^T)^T        contraction over the batch
    Accumulation is fp32; the parameter gradients come back in fp32."""

    @staticmethod
    def forward(ctx, x, weight, bias, relu):
        x2 = x.reshape(-1, x.shape[-1])
        xq, xs = quantize_mx(x2)
        wq, ws = quantize_mx(weight.detach())
        y = gemm_mxfp8(xq, xs, wq, ws, bias=bias, relu=relu, out_dtype=_torch.float32)
        ctx.save_for_backward(x2, weight, y if relu else None)
        ctx.relu, ctx.has_bias, ctx.lead = bool(relu), bias is not None, x.shape[:-1]
        return y.reshape(*x.shape[:-1], weight.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x2, weight, y = ctx.saved_tensors
        dy = dy.reshape(-1, dy.shape[-1]).float()
        if ctx.relu:
            dy = dy * (y > 0)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            gq, gs = quantize_mx(dy)
            wtq, wts = quantize_mx(weight.detach().t().contiguous())
            dx = gemm_mxfp8(gq, gs, wtq, wts, out_dtype=_torch.float32).reshape(*ctx.lead, weight.shape[1]).to(x2.dtype)
        if ctx.needs_input_grad[1]:
            gtq, gts = quantize_mx(dy.t().contiguous())
            xtq, xts = quantize_mx(x2.t().contiguous())
            dw = gemm_mxfp8(gtq, gts, xtq, xts, out_dtype=_torch.float32).to(weight.dtype)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = dy.sum(0)
        return dx, dw, db, None


def linear_fp8(x, weight, bias=None, relu=False):
    return Fp8LinearFn.apply(x, weight, bias, relu)
