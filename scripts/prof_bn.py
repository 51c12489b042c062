"""Isolated launches of the layer-1 sized fused BN/ReLU/pool kernels and conv1 kernels (ncu / timing)."""
import sys, torch
sys.path.insert(0, '.')
from coinstac_dinunet_b200.ops import vbm
dev = torch.device('cuda')
torch.manual_seed(0)
N, D, H, W, C = 8, 121, 145, 121, 16
x = torch.randn(N, D, H, W, device=dev)
w = torch.randn(16, 1, 3, 3, 3, device=dev) * 0.2
gamma, beta = torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev) * 0.1
y, stats = vbm.conv1_fwd(x, w)
mean, invstd = vbm.bn_finalize(stats, N * D * H * W, 1e-5, 0.1)
p = vbm.bn_relu_pool_fwd(y, mean, invstd, gamma, beta)
dp = torch.randn_like(p)

def timeit(name, fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    print(f'{name:28s} {e0.elapsed_time(e1) / n * 1e3:8.1f} us')

if len(sys.argv) > 1 and sys.argv[1] == 'prof_t':
    xp = vbm.conv1_pad_input(x)
    vbm.conv1_fwd(x, w, impl='toeplitz', xp=xp)
    torch.cuda.synchronize()
elif len(sys.argv) > 1 and sys.argv[1] == 'prof':
    dy, dg, db = vbm.bn_relu_pool_bwd(y, dp, mean, invstd, gamma, beta)
    vbm.conv1_wgrad(dy, x)
    vbm.conv1_fwd(x, w)
    torch.cuda.synchronize()
else:
    timeit('conv1_fwd[cuda]', lambda: vbm.conv1_fwd(x, w, impl='cuda'))
    timeit('conv1_fwd[tc]', lambda: vbm.conv1_fwd(x, w, impl='tc'))
    timeit('conv1_pad_input', lambda: vbm.conv1_pad_input(x))
    xp = vbm.conv1_pad_input(x)
    timeit('conv1_fwd[toeplitz] (no pad)', lambda: vbm.conv1_fwd(x, w, impl='toeplitz', xp=xp))
    if len(sys.argv) > 1 and sys.argv[1] == 'fwdonly': sys.exit(0)
    timeit('bn_relu_pool_fwd', lambda: vbm.bn_relu_pool_fwd(y, mean, invstd, gamma, beta))
    timeit('bn_relu_pool_bwd (A+B)', lambda: vbm.bn_relu_pool_bwd(y, dp, mean, invstd, gamma, beta))
    dy, _, _ = vbm.bn_relu_pool_bwd(y, dp, mean, invstd, gamma, beta)
    timeit('conv1_wgrad[cuda]', lambda: vbm.conv1_wgrad(dy, x, impl='cuda'))
    timeit('conv1_wgrad[tc]', lambda: vbm.conv1_wgrad(dy, x, impl='tc'))
    # roofline reference: a plain copy of the same bytes
    buf = torch.empty_like(y)
    timeit('copy y (544 MB r + 544 MB w)', lambda: buf.copy_(y))
