"""Where does run-to-run noise enter the native forward?  Repeats each stage on identical inputs."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from coinstac_dinunet_b200.ops import vbm, conv3d as c3

dev = torch.device('cuda', 0)
torch.manual_seed(0)


def spread(ts):
    ref = ts[0].double()
    return max(float(((t.double() - ref).abs() / (ref.abs() + 1e-30)).max()) for t in ts[1:])


def nbits(ts):
    return max(float((t.float() != ts[0].float()).float().mean()) for t in ts[1:])


for shape, N in (((33, 34, 35), 4), ((121, 145, 121), 8)):
    D, H, W = shape
    x = torch.randn(N, D, H, W, device=dev)
    w1 = torch.randn(16, 1, 3, 3, 3, device=dev) * 0.27
    xp = vbm.conv1_pad_input_hd(x)
    st = [vbm.conv1_fused_stats(xp, w1, (N, D, H, W)).clone() for _ in range(5)]
    print(f'shape {shape}: conv1 fused stats max rel spread over 5 runs: {spread(st):.3e}')
    # fp64 oracle of the sums on bf16-rounded operands
    y = torch.nn.functional.conv3d(x.bfloat16().double().unsqueeze(1), w1.bfloat16().double(), padding=1)
    s_ref = torch.cat([y.sum((0, 2, 3, 4)), (y * y).sum((0, 2, 3, 4))])
    print(f'   vs fp64 oracle: rel err sum {float(((st[0][:16].double() - s_ref[:16]).abs() / s_ref[:16].abs()).max()):.3e} '
          f'sumsq {float(((st[0][16:].double() - s_ref[16:]).abs() / s_ref[16:].abs()).max()):.3e}')
    n = x.numel()
    mean = (s_ref[:16] / n).float(); var = (s_ref[16:] / n - (s_ref[:16] / n) ** 2).float()
    invstd = (var + 1e-5).rsqrt()
    g, b = torch.ones(16, device=dev), torch.zeros(16, device=dev)
    ps = [vbm.conv1_fused_pool(xp, w1, mean, invstd, g, b, (N, D, H, W)) for _ in range(4)]
    print(f'   pool output mismatch fraction run-to-run: {nbits([p for p, _ in ps]):.3e}  code: {nbits([c for _, c in ps]):.3e}')
    p = ps[0][0]
    for cin, cout in ((16, 32), (32, 64), (64, 128), (128, 256)):
        if min(p.shape[1:4]) < 2:
            break
        w = torch.randn(cout, cin, 3, 3, 3, device=dev) * (27 * cin) ** -0.5
        outs = [c3.conv3d_igemm_fwd(p, w, want_stats=True) for _ in range(4)]
        ys = [o[0] for o in outs]
        sts = [o[1] for o in outs if o[1] is not None]
        msg = f'   conv {cin}->{cout} on {tuple(p.shape)} [{c3.last_impl}]: y mismatch fraction {nbits(ys):.3e}'
        if len(sts) > 1:
            yf = ys[0].double().reshape(-1, cout)
            s_ref = torch.cat([yf.sum(0), (yf * yf).sum(0)])
            msg += f'  stats spread {spread(sts):.3e}  vs fp64 of stored y: {float(((sts[0].double() - s_ref).abs() / s_ref.abs()).max()):.3e}'
        else:
            sb = [vbm.bn_stats(ys[0]).clone() for _ in range(4)]
            yf = ys[0].double().reshape(-1, cout)
            s_ref = torch.cat([yf.sum(0), (yf * yf).sum(0)])
            msg += f'  bn_stats spread {spread(sb):.3e} vs fp64: {float(((sb[0].double() - s_ref).abs() / s_ref.abs()).max()):.3e}'
        print(msg)
        yy = ys[0]
        cnt = yy.numel() // cout
        m = yy.float().reshape(-1, cout).mean(0); v = yy.float().reshape(-1, cout).var(0, unbiased=False)
        pp = [vbm.bn_relu_pool_fwd(yy, m, (v + 1e-5).rsqrt(), torch.ones(cout, device=dev), torch.zeros(cout, device=dev)) for _ in range(3)]
        print(f'      bn_relu_pool mismatch fraction {nbits(pp):.3e}')
        # backward pieces on fixed inputs
        dy = torch.randn_like(yy)
        dxs, dws = [], []
        for _ in range(3):
            dx, dw = c3.conv3d_igemm_bwd(dy, p, w, need_dx=True)
            dxs.append(dx); dws.append(dw.clone())
        print(f'      dgrad mismatch fraction {nbits(dxs):.3e}  wgrad rel spread {max(float((d - dws[0]).norm() / dws[0].norm()) for d in dws[1:]):.3e}')
        p = pp[0]
