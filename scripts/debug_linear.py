import torch, sys
sys.path.insert(0, '.')
from coinstac_dinunet_b200.ops.linear import gemm_tn, LinearFn
dev = 'cuda'
torch.manual_seed(0)
M, N, K = 24, 72, 200
x = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev) * 0.07; b = torch.randn(N, device=dev) * 0.1
xb, wb = x.bfloat16(), w.bfloat16()
dy = torch.randn(M, N, device=dev).bfloat16()
ref_dx = dy.float() @ wb.float()
got = gemm_tn(dy, wb.t().contiguous(), out_dtype=torch.bfloat16).float()
err = (got - ref_dx).abs()
print('dx rel', float((got-ref_dx).norm()/ref_dx.norm()), 'max abs', float(err.max()), 'argmax', divmod(int(err.argmax()), K))
print('col err', (err.max(0).values > 0.05).nonzero().flatten().tolist()[:40])
print('row err', (err.max(1).values > 0.05).nonzero().flatten().tolist()[:40])
ref_dw = dy.float().t() @ xb.float()
gotw = gemm_tn(dy.t().contiguous(), xb.t().contiguous(), out_dtype=torch.float32)
print('dw rel', float((gotw-ref_dw).norm()/ref_dw.norm()))
y = gemm_tn(xb, wb, bias=b, relu=True)
ref_y = (xb.float() @ wb.float().t() + b).relu()
print('y rel', float((y.float()-ref_y).norm()/ref_y.norm()))
