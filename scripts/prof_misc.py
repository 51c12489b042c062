"""Isolated launches for ncu: tcgen05 GEMM (head of VBMNet), fused local optimizer step, halo conv after hoisting."""
import sys, torch
sys.path.insert(0, '.')
from coinstac_dinunet_b200.ops.linear import gemm_tn
from coinstac_dinunet_b200.parallel.arena import DistArena
dev = torch.device('cuda')
torch.manual_seed(0)
a = torch.randn(4096, 4096, device=dev).bfloat16(); b = torch.randn(4096, 4096, device=dev).bfloat16()
for _ in range(3): c = gemm_tn(a, b, split_k=1)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): c = gemm_tn(a, b, split_k=1)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f'gemm 4096^3 bf16: {ms*1e3:.1f} us = {2*4096**3/ms/1e9:.1f} TFLOP/s (measured cuBLAS peak 1674.9)')
want = a.float() @ b.float().t()
print('rel err', float((c.float() - want).norm() / want.norm()))
m = torch.nn.Linear(4096, 4096 * 8).to(dev)       # 134M params: 537 MB of fp32 gradients
o = torch.optim.Adam(m.parameters(), lr=1e-3)
ar = DistArena(m, o, device=dev, backend='nvlink')
for _ in range(3): ar.flat_grad.normal_(); ar.reduce_and_step()
torch.cuda.synchronize()
e0.record()
for _ in range(5): ar.reduce_and_step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
nb = ar.numel * 4
print(f'fused Adam S=1: {ms*1e3:.1f} us for {nb/1e6:.0f} MB grads; {7*nb/ms/1e6:.0f} GB/s of 6578.7 (4 reads + 3 writes + zero = 8 x 4 B per param)')
