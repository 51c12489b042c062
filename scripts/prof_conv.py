"""Isolated launches of the layer-2 convolution kernels (batch 8 of 16ch x 60x72x60) for ncu captures and timing."""
import sys, torch
sys.path.insert(0, '.')
from coinstac_dinunet_b200.ops.conv3d import conv3d_igemm_fwd, conv3d_igemm_bwd
from coinstac_dinunet_b200.ops import conv3d as c3
dev = torch.device('cuda')
torch.manual_seed(0)
N, D, H, W = 8, 60, 72, 60
x = torch.randn(N, D, H, W, 16, device=dev).bfloat16()
w = torch.randn(32, 16, 3, 3, 3, device=dev) * 0.05
dy = torch.randn(N, D, H, W, 32, device=dev).bfloat16()
which = sys.argv[1] if len(sys.argv) > 1 else 'time'
if which == 'time':
    for name, fn in (('fprop16->32', lambda: conv3d_igemm_fwd(x, w)), ('bwd(dgrad+wgrad)', lambda: conv3d_igemm_bwd(dy, x, w))):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): fn()
        e1.record(); torch.cuda.synchronize()
        print(name, c3.last_impl, round(e0.elapsed_time(e1) / 10 * 1e3, 1), 'us')
else:
    for _ in range(3):
        y = conv3d_igemm_fwd(x, w)
        dx, dw = conv3d_igemm_bwd(dy, x, w)
    torch.cuda.synchronize()
