"""Conv3d (k3, p1, s1) on the tcgen05 implicit-GEMM kernel (``csrc/conv3d_tcgen05.cu``).

Weight layouts fed to the kernel (bf16, K padded to a multiple of 64 with zeros):
    fprop   Wk[co, tap*Cin + ci]   = W[co, ci, kd, kh, kw]
    dgrad   Wd[ci, tap'*Cout + co] = W[co, ci, 2-kd', 2-kh', 2-kw']      (dx = conv(dy, Wd))
"""
import os as _os

import torch as _torch

from . import native as _nat

BF16 = _torch.bfloat16
_SUPPORTED = {(16, 32), (32, 64), (64, 128), (128, 256), (32, 16), (64, 32), (128, 64), (256, 128),
              (16, 16), (32, 32), (64, 64), (128, 128)}


def supported(cin, cout):
    return (cin, cout) in _SUPPORTED


def _bump(n=1):
    from . import _count_launch
    _count_launch(n)


def _pad_k(w2d):
    k = w2d.shape[1]
    kpad = (k + 63) // 64 * 64
    if kpad != k:
        w2d = _torch.nn.functional.pad(w2d, (0, kpad - k))
    return w2d.contiguous(), kpad


def pack_fprop_weight(weight):
    cout, cin = weight.shape[:2]
    return _pad_k(weight.detach().permute(0, 2, 3, 4, 1).reshape(cout, 27 * cin).to(BF16))


def pack_dgrad_weight(weight):
    cout, cin = weight.shape[:2]
    return _pad_k(weight.detach().flip(2, 3, 4).permute(1, 2, 3, 4, 0).reshape(cin, 27 * cout).to(BF16))


def conv_impl():
    """'auto' (default): halo kernel (v3, every input voxel crosses L2->SM once) where it applies, else the
    TMA-box kernel (v2).  'halo' / 'tma' / 'gather' force one implementation ('gather' = v1, cp.async)."""
    import os
    return os.environ.get('COINN_CONV_IMPL', 'auto')


#: which implementation served the last call (tests / profiling)
stats_done = False
last_impl = None


_pack_buffers = {}
_pack_stamp = {}      # which tensor object / version the cached pack belongs to


def pack_weights(weight):
    """Both GEMM operand layouts of a conv weight in ONE launch (``vbm_fused.cu::pack_conv_weights_kernel``).
    The bf16 buffers persist per parameter (their zero padding is written once); contents are refreshed on every
    call because the optimizer changes the weights every step.  Returns (wf, kf, wd, kd)."""
    cout, cin = weight.shape[:2]
    kf, kd = (27 * cin + 63) // 64 * 64, (27 * cout + 63) // 64 * 64
    key = (weight.data_ptr(), cout, cin, weight.device)
    bufs = _pack_buffers.get(key)
    if bufs is None:
        bufs = (_torch.zeros((cout, kf), dtype=BF16, device=weight.device),
                _torch.zeros((cin, kd), dtype=BF16, device=weight.device))
        _pack_buffers[key] = bufs
    w32 = weight.detach()
    if w32.dtype != _torch.float32 or not w32.is_contiguous():
        w32 = w32.float().contiguous()
    _nat.check(_nat.lib().coinn_pack_conv_weights(w32.data_ptr(), bufs[0].data_ptr(), bufs[1].data_ptr(), cout, cin, kf, kd,
                                                  _nat.stream_ptr(weight.device)), 'coinn_pack_conv_weights')
    _bump()
    _pack_stamp[key] = (id(weight), weight._version)
    return bufs[0], kf, bufs[1], kd


FP8_PAIRS = {(32, 64), (64, 128), (128, 256), (64, 32), (128, 64), (256, 128), (32, 32), (64, 64), (128, 128)}


def _quantize_rows(x2d, K, box):
    """x2d: [R, >= K] (row stride = x2d.stride(0)), box = channel-box width in elements (32 / 64 / 128)."""
    assert x2d.stride(1) == 1 and K % box == 0 and box in (32, 64, 128)
    if x2d.dtype not in (BF16, _torch.float32):
        x2d = x2d.float()
    R, G = x2d.shape[0], box // 32
    q = _torch.empty((R, K), dtype=_torch.uint8, device=x2d.device)
    sf = _torch.full((R, K // box), 0x7f7f7f7f, dtype=_torch.int32, device=x2d.device)
    _nat.check(_nat.lib().coinn_quantize_mx_grouped(x2d.data_ptr(), 1 if x2d.dtype == BF16 else 0, x2d.stride(0), q.data_ptr(),
                                                    sf.data_ptr(), R, K, G, _nat.stream_ptr(x2d.device)), 'quantize_mx_grouped')
    _bump(2)
    return q, sf


def _igemm_fp8(x, wk, cin, cout):
    """MX-FP8 implicit-GEMM conv (``csrc/conv3d_mxfp8.cu``): x [N,D,H,W,cin] bf16, wk [cout, >= 27*cin] bf16 packed
    (k = tap * cin + ci).  Activations and weights are quantised per 32 input channels (ue8m0 scales) right here."""
    global last_impl
    N, D, H, W, _ = x.shape
    box = min(cin, 128)
    xq, xsf = _quantize_rows(x.reshape(-1, cin), cin, box)
    wq, wsf = _quantize_rows(wk, 27 * cin, box)
    y = _torch.empty((N, D, H, W, cout), dtype=BF16, device=x.device)
    code = _nat.lib().coinn_conv3d_mxfp8(xq.data_ptr(), xsf.data_ptr(), wq.data_ptr(), wsf.data_ptr(), y.data_ptr(),
                                         N, D, H, W, cin, cout, _nat.stream_ptr(x.device))
    _nat.check(code, f'conv3d_mxfp8({cin}->{cout})')
    last_impl = 'mxfp8'
    _bump()
    return y


def _igemm(x, wk, kpad, cout, impl=None, stats=None):
    """``stats``: optional zeroed fp32 [2*cout]; filled with the BatchNorm sums of y by the halo kernel's epilogue
    (returns False in ``stats_done`` when the kernel that ran cannot do it)."""
    global last_impl, stats_done
    N, D, H, W, cin = x.shape
    y = _torch.empty((N, D, H, W, cout), dtype=BF16, device=x.device)
    impl = impl or conv_impl()
    lib = _nat.lib()
    args = (x.data_ptr(), wk.data_ptr(), y.data_ptr(), N, D, H, W, cin, cout, kpad, _nat.stream_ptr(x.device))
    code = -1
    stats_done = False
    if impl in ('auto', 'halo', 'halo2'):
        fullpix = 1 if impl == 'halo2' or (impl == 'auto' and _os.environ.get('COINN_HALO_FULLPIX', '1') == '1') else 0
        want = stats is not None and cout <= 64
        code = lib.coinn_conv3d_halo_stats(args[0], args[1], args[2], stats.data_ptr() if want else None, *args[3:-1], fullpix, args[-1])
        last_impl = 'halo2' if fullpix else 'halo'
        stats_done = want and code == 0
    if code == -1 and impl != 'gather':
        code, last_impl = lib.coinn_conv3d_tma(*args), 'tma'
    if code == -1:
        code, last_impl = lib.coinn_conv3d_igemm(*args), 'gather'
    _nat.check(code, f'conv3d[{last_impl}]({cin}->{cout})')
    _bump()
    return y


def conv3d_igemm_fwd(x, weight, want_stats=False, stats_out=None, fp8=False):
    """x: [N,D,H,W,Cin] bf16 contiguous; weight: [Cout,Cin,3,3,3] -> [N,D,H,W,Cout] bf16.
    ``fp8``: MX-FP8 block-scaled kernel where it is instantiated (C_in >= 32), bf16 otherwise."""
    cout, cin = weight.shape[:2]
    if not supported(cin, cout):
        raise ImportError(f'no tcgen05 conv instantiation for {cin}->{cout}')
    wk, kpad, _, _ = pack_weights(weight)
    if fp8 and (cin, cout) in FP8_PAIRS:
        y = _igemm_fp8(x.contiguous(), wk, cin, cout)
        return (y, None) if want_stats else y
    if not want_stats:
        return _igemm(x.contiguous(), wk, kpad, cout)
    stats = stats_out if stats_out is not None else _torch.zeros(2 * cout, dtype=_torch.float32, device=x.device)
    y = _igemm(x.contiguous(), wk, kpad, cout, stats=stats)
    return y, (stats if stats_done else None)


def conv3d_igemm_bwd(dy, x, weight, need_dx=True, raw_dw=None, fp8=False):
    cout, cin = weight.shape[:2]
    if not supported(cout, cin):
        raise ImportError(f'no tcgen05 conv instantiation for dgrad {cout}->{cin}')
    dx = None
    if need_dx:
        key = (weight.data_ptr(), cout, cin, weight.device)
        if _pack_stamp.get(key) == (id(weight), weight._version):   # packed by this step's forward, weights unchanged since
            wd, kpad = _pack_buffers[key][1], (27 * cout + 63) // 64 * 64
        else:
            _, _, wd, kpad = pack_weights(weight)
        if fp8 and (cout, cin) in FP8_PAIRS:          # dgrad = conv of dy (C = cout) with the flipped / transposed weights
            dx = _igemm_fp8(dy.contiguous(), wd, cout, cin)
        else:
            dx = _igemm(dy.contiguous(), wd, kpad, cin)
    try:
        from .conv3d_wgrad import conv3d_wgrad
        dw = conv3d_wgrad(dy, x, raw_out=raw_dw)
    except ImportError:
        # interim: cuDNN weight gradient (library) until the tcgen05 wgrad kernel lands
        w = weight.detach().to(BF16).contiguous(memory_format=_torch.channels_last_3d)
        _, dw, _ = _torch.ops.aten.convolution_backward(
            dy.permute(0, 4, 1, 2, 3), x.permute(0, 4, 1, 2, 3), w, None, [1, 1, 1], [1, 1, 1], [1, 1, 1], False,
            [0, 0, 0], 1, [False, True, False])
        dw = dw.float()
    return dx, dw
