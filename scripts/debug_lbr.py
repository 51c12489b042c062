import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from coinstac_dinunet_b200.ops.linear import SmallLinearFn, SmallLinearBnReluFn
dev = torch.device('cuda', 0)
torch.manual_seed(0)
print('tf32 matmul allowed:', torch.backends.cuda.matmul.allow_tf32, torch.get_float32_matmul_precision())
for M, K, N in ((2, 9, 5), (16, 66, 256)):
    lin, bn = torch.nn.Linear(K, N).to(dev), torch.nn.BatchNorm1d(N).to(dev)
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5); bn.bias.normal_(0, 0.3)
    x = torch.randn(M, K, device=dev)
    y_t = torch.nn.functional.linear(x, lin.weight, lin.bias)
    y_d = (x.double() @ lin.weight.double().t() + lin.bias.double()).float()
    y_k = SmallLinearFn.apply(x, lin.weight, lin.bias, False)
    print(M, K, N, 'linear: kernel-vs-fp64', float((y_k - y_d).abs().max()), 'torch-vs-fp64', float((y_t - y_d).abs().max()))
    rm, rv, nbt = torch.zeros(N, device=dev), torch.ones(N, device=dev), torch.zeros((), dtype=torch.long, device=dev)
    z = SmallLinearBnReluFn.apply(x, lin.weight, lin.bias, bn.weight, bn.bias, rm, rv, nbt, 1e-5, 0.1, True, False)
    mean, var = y_d.mean(0), y_d.var(0, unbiased=False)
    z_m = bn.weight * (y_d - mean) / (var + 1e-5).sqrt() + bn.bias
    z_t = torch.nn.functional.batch_norm(y_t, None, None, bn.weight, bn.bias, True, 0.1, 1e-5)
    print('   fused-vs-manual', float((z - z_m).abs().max()), 'torchbn-vs-manual', float((z_t - z_m).abs().max()))
    print('   col0 fused', z[:, 0].tolist()[:4], 'manual', z_m[:, 0].tolist()[:4], 'beta0', float(bn.bias[0]), 'gamma0', float(bn.weight[0]))
    print('   running_mean err', float((rm - 0.1 * mean).abs().max()), 'nbt', int(nbt))
