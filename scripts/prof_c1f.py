"""Layer-1 (conv1 + BN + ReLU + pool) forward/backward: fused-recompute kernels vs the stored-y pipeline."""
import sys, torch
sys.path.insert(0, '.')
from coinstac_dinunet_b200.ops import vbm
dev = torch.device('cuda')
torch.manual_seed(0)
N, D, H, W = 8, 121, 145, 121
shape = (N, D, H, W)
x = torch.randn(N, D, H, W, device=dev).bfloat16()
w = torch.randn(16, 1, 3, 3, 3, device=dev) * 0.2
gamma, beta = torch.rand(16, device=dev) + 0.5, torch.randn(16, device=dev) * 0.1

def timeit(name, fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    print(f'{name:36s} {e0.elapsed_time(e1) / n * 1e3:8.1f} us', flush=True)

xp = vbm.conv1_pad_input_hd(x)
stats = vbm.conv1_fused_stats(xp, w, shape)
mean, invstd = vbm.bn_finalize(stats, x.numel(), 1e-5, 0.1)
p, code = vbm.conv1_fused_pool(xp, w, mean, invstd, gamma, beta, shape)
dp = torch.randn_like(p)
if len(sys.argv) > 1 and sys.argv[1] == 'prof':
    vbm.conv1_fused_bwd(xp, w, mean, invstd, gamma, beta, p, code, dp, shape)
    torch.cuda.synchronize()
    sys.exit(0)
timeit('pad_input_hd (bf16 in)', lambda: vbm.conv1_pad_input_hd(x))
timeit('fused stats', lambda: vbm.conv1_fused_stats(xp, w, shape))
timeit('fused conv+bn+relu+pool', lambda: vbm.conv1_fused_pool(xp, w, mean, invstd, gamma, beta, shape))
timeit('fused bwd (pooled stats + wgrad)', lambda: vbm.conv1_fused_bwd(xp, w, mean, invstd, gamma, beta, p, code, dp, shape))
# the stored-y pipeline it replaces
xf = x.float()
y, st = vbm.conv1_fwd(xf, w, impl='toeplitz')
timeit('[old] conv1_fwd toeplitz (+pad)', lambda: vbm.conv1_fwd(xf, w, impl='toeplitz'))
timeit('[old] bn_relu_pool_fwd', lambda: vbm.bn_relu_pool_fwd(y, mean, invstd, gamma, beta))
po = vbm.bn_relu_pool_fwd(y, mean, invstd, gamma, beta)
timeit('[old] bn_relu_pool_bwd', lambda: vbm.bn_relu_pool_bwd(y, dp, mean, invstd, gamma, beta, p=po))
dy, _, _ = vbm.bn_relu_pool_bwd(y, dp, mean, invstd, gamma, beta, p=po)
timeit('[old] conv1_wgrad tc', lambda: vbm.conv1_wgrad(dy, xf, impl='tc'))
