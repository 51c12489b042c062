"""Launch the round-2 kernels once at realistic shapes (driver for `ncu -k ... -c 1`)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from coinstac_dinunet_b200.ops import conv3d as c3
from coinstac_dinunet_b200.ops.fp8 import gemm_mxfp8, quantize_mx
from coinstac_dinunet_b200.ops.linear import linear_bn_relu
dev = torch.device('cuda', 0)
torch.manual_seed(0)
which = sys.argv[1]
if which == 'conv_fp8':            # block 3 of the benchmark: 8 x 30x36x30, 32 -> 64
    x = torch.randn(8, 30, 36, 30, 32, device=dev).bfloat16(); w = torch.randn(64, 32, 3, 3, 3, device=dev) * 0.03
    for _ in range(3):
        c3.conv3d_igemm_fwd(x, w, fp8=True)
    c3.conv3d_igemm_fwd(x, w)      # the bf16 halo kernel on the same shape, for comparison
elif which == 'gemm_fp8':
    a = torch.randn(4096, 4096, device=dev); b = torch.randn(4096, 4096, device=dev)
    aq, asf = quantize_mx(a); bq, bsf = quantize_mx(b)
    for _ in range(3):
        gemm_mxfp8(aq, asf, bq, bsf, split_k=1, out_dtype=torch.bfloat16)
elif which == 'lbr':
    lin, bn = torch.nn.Linear(66, 256).to(dev), torch.nn.BatchNorm1d(256).to(dev)
    x = torch.randn(16, 66, device=dev, requires_grad=True)
    for _ in range(3):
        linear_bn_relu(x, lin, bn).sum().backward()
torch.cuda.synchronize()
