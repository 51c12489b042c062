"""Conv3d weight gradient on the tcgen05 implicit-GEMM kernel (``csrc/conv3d_wgrad_tcgen05.cu``)."""
import os as _os

import torch as _torch

from . import native as _nat

last_impl = None
_SUPPORTED = {(16, 32), (32, 64), (64, 128), (128, 256), (32, 32), (64, 64)}


def conv3d_wgrad(dy, x, raw_out=None):
    """dy: [N,D,H,W,Cout] bf16, x: [N,D,H,W,Cin] bf16 -> dW [Cout,Cin,3,3,3] fp32.
    ``raw_out``: fp32 [27*Cin, Cout] accumulator in the kernels' own (tap, ci, co) layout; the gradient is ADDED to it and
    nothing is returned (``coinn_conv_block_grad_finalize`` folds it into ``weight.grad`` later)."""
    N, D, H, W, cin = x.shape
    cout = dy.shape[-1]
    if (cin, cout) not in _SUPPORTED:
        raise ImportError(f'no tcgen05 wgrad instantiation for {cin}->{cout}')
    dwt = raw_out if raw_out is not None else _torch.zeros((27 * cin, cout), dtype=_torch.float32, device=x.device)
    impl = _os.environ.get('COINN_WGRAD_IMPL', 'auto')      # auto: halo kernel, else the per-tap kernel, else the gather kernel
    args = (x.contiguous().data_ptr(), dy.contiguous().data_ptr(), dwt.data_ptr(), N, D, H, W, cin, cout,
            _nat.stream_ptr(x.device))
    global last_impl
    code = -1
    if impl in ('auto', 'halo'):
        code, last_impl = _nat.lib().coinn_conv3d_wgrad_halo(*args), 'halo'
    if code == -1 and impl in ('auto', 'tap'):          # C_in >= 64: one TMA-box GEMM per filter tap (conv3d_wgrad_tap.cu)
        code, last_impl = _nat.lib().coinn_conv3d_wgrad_tap(*args), 'tap'
    if code == -1:
        code, last_impl = _nat.lib().coinn_conv3d_wgrad(*args), 'gather'
    _nat.check(code, f'conv3d_wgrad[{last_impl}]({cin}->{cout})')
    from . import _count_launch
    _count_launch()
    if raw_out is not None:
        return None
    # dwt[(kd,kh,kw,ci), co] -> [co, ci, kd, kh, kw]
    return dwt.view(3, 3, 3, cin, cout).permute(4, 3, 0, 1, 2).contiguous()
