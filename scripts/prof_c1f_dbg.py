import sys, torch
sys.path.insert(0, '.')
from coinstac_dinunet_b200.ops import vbm
dev = torch.device('cuda'); torch.manual_seed(0)
shape = (8, 121, 145, 121)
x = torch.randn(*shape, device=dev).bfloat16(); w = torch.randn(16, 1, 3, 3, 3, device=dev) * 0.2
xp = vbm.conv1_pad_input_hd(x)
for _ in range(3): vbm.conv1_fused_stats(xp, w, shape)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): vbm.conv1_fused_stats(xp, w, shape)
e1.record(); torch.cuda.synchronize()
print('stats', e0.elapsed_time(e1) / 10 * 1e3, 'us (incl. 1 memset)', flush=True)
