"""Numerics of the hand-written kernels against plain PyTorch fp32 oracles (single B200)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def dev():
    from coinstac_dinunet_b200 import ops
    assert ops.native_available(), 'native kernel library must be loaded on a GPU box'
    return torch.device('cuda:0')


def test_count_binary_matches_reference_coding(dev):
    from coinstac_dinunet_b200 import ops
    g = torch.Generator(device='cpu').manual_seed(1)
    for n, dt in ((1, torch.int64), (1000, torch.int64), (123457, torch.uint8), (4096, torch.int32)):
        pred = torch.randint(0, 2, (n,), generator=g).to(dt)
        true = torch.randint(0, 2, (n,), generator=g).to(dt)
        if dt == torch.uint8:
            pred, true = pred * 255, true * 255          # 8-bit masks: 255 counts as 1
        counter = torch.zeros(4, dtype=torch.int64, device=dev)
        ops.count_binary(pred.to(dev), true.to(dev), counter)
        p, t = (pred != 0).long(), (true != 0).long()
        want = torch.bincount(2 * t + p, minlength=4)
        assert counter.cpu().tolist() == want.tolist()


def test_prf1a_device_path_equals_cpu(dev):
    from coinstac_dinunet_b200.metrics import Prf1a
    g = torch.Generator().manual_seed(3)
    a, b = Prf1a(), Prf1a()
    for _ in range(5):
        pred, true = torch.randint(0, 2, (257,), generator=g), torch.randint(0, 2, (257,), generator=g)
        a.add(pred, true)
        b.add(pred.to(dev), true.to(dev))
    assert a.get() == b.get() and (a.tp, a.fp, a.tn, a.fn) == (b.tp, b.fp, b.tn, b.fn)


def test_count_confusion(dev):
    from coinstac_dinunet_b200 import ops
    g = torch.Generator().manual_seed(2)
    for C, n in ((3, 10), (7, 5000), (10, 100003)):
        pred, true = torch.randint(0, C, (n,), generator=g), torch.randint(0, C, (n,), generator=g)
        mat = torch.zeros(C, C, dtype=torch.int64, device=dev)
        ops.count_confusion(pred.to(dev), true.to(dev), mat)
        want = torch.bincount(pred * C + true, minlength=C * C).view(C, C)
        assert torch.equal(mat.cpu(), want)


@pytest.mark.parametrize('dtype,tol', [(torch.float32, 1e-5), (torch.bfloat16, 2e-2), (torch.float16, 2e-3)])
def test_softmax_nll_forward_backward(dev, dtype, tol):
    from coinstac_dinunet_b200 import ops
    torch.manual_seed(0)
    for n, c in ((1, 2), (16, 2), (33, 5), (257, 100)):
        logits = (torch.randn(n, c, device=dev) * 3).to(dtype).requires_grad_(True)
        labels = torch.randint(0, c, (n,), device=dev)
        loss, pred = ops.softmax_nll(logits, labels)
        ref_in = logits.detach().float().requires_grad_(True)
        ref = torch.nn.functional.nll_loss(torch.log_softmax(ref_in, 1), labels)
        assert torch.allclose(loss, ref, atol=tol, rtol=tol)
        assert torch.equal(pred, ref_in.argmax(1))
        (loss * 2.5).backward()
        (ref * 2.5).backward()
        assert torch.allclose(logits.grad.float(), ref_in.grad, atol=tol, rtol=tol)


def test_orthogonalize_matches_gram_schmidt(dev):
    from coinstac_dinunet_b200 import ops
    from coinstac_dinunet_b200.distrib.powersgd import _orthogonalize
    torch.manual_seed(0)
    for m, r in ((5, 1), (100, 4), (4097, 8), (300, 32)):
        a = torch.randn(m, r, device=dev)
        want = _orthogonalize(a.clone())
        got = ops.orthogonalize_(a.clone())
        assert torch.allclose(got, want, atol=2e-4, rtol=1e-3)
        eye = got.t() @ got
        assert torch.allclose(eye, torch.eye(r, device=dev), atol=1e-3)


def _mlp(dev):
    torch.manual_seed(5)
    return torch.nn.Sequential(torch.nn.Linear(37, 53), torch.nn.ReLU(), torch.nn.Linear(53, 3)).to(dev)


@pytest.mark.parametrize('make_opt', [
    lambda p: torch.optim.Adam(p, lr=1e-2),
    lambda p: torch.optim.Adam(p, lr=1e-2, weight_decay=0.1),
    lambda p: torch.optim.AdamW(p, lr=1e-2, weight_decay=0.1),
    lambda p: torch.optim.SGD(p, lr=0.1),
    lambda p: torch.optim.SGD(p, lr=0.1, momentum=0.9, nesterov=True, weight_decay=0.01),
])
def test_fused_local_step_matches_torch_optimizer(dev, make_opt):
    """S == 1 path of fused_reduce_opt.cu vs torch.optim on the same gradients, 6 steps."""
    from coinstac_dinunet_b200.parallel.arena import DistArena
    ref, ours = _mlp(dev), _mlp(dev)
    ref_opt, our_opt = make_opt(ref.parameters()), make_opt(ours.parameters())
    arena = DistArena(ours, our_opt, device=dev, backend='nvlink')
    assert arena.backend == 'nvlink'
    for step in range(6):
        x = torch.randn(19, 37, device=dev)
        for net in (ref, ours):
            net(x).square().mean().backward()
        ref_opt.step(); ref_opt.zero_grad()
        arena.reduce_and_step()
        assert float(arena.flat_grad.abs().max()) == 0.0
        for a, b in zip(ref.parameters(), ours.parameters()):
            assert torch.allclose(a, b, atol=2e-6, rtol=2e-5), step


def test_arena_checkpoint_state_roundtrip(dev, tmp_path):
    from coinstac_dinunet_b200.parallel.arena import DistArena
    net = _mlp(dev)
    opt = torch.optim.Adam(net.parameters(), lr=1e-2)
    arena = DistArena(net, opt, device=dev, backend='nvlink')
    for _ in range(3):
        net(torch.randn(4, 37, device=dev)).sum().backward()
        arena.reduce_and_step()
    arena.gather_state()
    sd = opt.state_dict()
    assert sd['state'][0]['step'] == 3 and sd['state'][0]['exp_avg'].abs().sum() > 0
    torch.save({'m': net.state_dict(), 'o': sd}, tmp_path / 'c.pt')
    net2 = _mlp(dev)
    opt2 = torch.optim.Adam(net2.parameters(), lr=1e-2)
    chk = torch.load(tmp_path / 'c.pt', weights_only=False)
    net2.load_state_dict(chk['m']); opt2.load_state_dict(chk['o'])
    arena2 = DistArena(net2, opt2, device=dev, backend='nvlink')
    assert int(arena2.step_count.item()) == 3
    assert torch.allclose(arena2.m, arena.m) and torch.allclose(arena2.v, arena.v)
    assert torch.equal(arena2.flat_param, arena.flat_param)


@pytest.mark.parametrize('M,N,K', [(128, 128, 64), (8, 256, 9216), (16, 2, 64), (300, 200, 136), (1, 16, 8),
                                   (257, 129, 72), (16, 256, 66), (1024, 512, 1024)])
def test_tcgen05_gemm_matches_fp32(dev, M, N, K):
    from coinstac_dinunet_b200.ops.linear import gemm_tn
    torch.manual_seed(M * 7 + N)
    a = torch.randn(M, K, device=dev).to(torch.bfloat16)
    b = torch.randn(N, K, device=dev).to(torch.bfloat16)
    bias = torch.randn(N, device=dev)
    want = a.float() @ b.float().t()
    for split in (1, None, 3):
        got = gemm_tn(a, b, out_dtype=torch.float32, split_k=split)
        assert torch.allclose(got, want, atol=1e-2 * K ** 0.5, rtol=1e-2), (split, (got - want).abs().max())
    got = gemm_tn(a, b, bias=bias, relu=True, out_dtype=torch.bfloat16, split_k=1)
    ref = (want + bias).relu()
    assert torch.allclose(got.float(), ref, atol=0.05 * K ** 0.5, rtol=2e-2)


def _rel(got, want):
    return float((got.float() - want.float()).norm() / want.float().norm().clamp_min(1e-12))


def test_b200_linear_autograd(dev):
    from coinstac_dinunet_b200.ops.linear import B200Linear
    torch.manual_seed(0)
    ref = torch.nn.Linear(200, 72).to(dev)
    ours = B200Linear.from_linear(torch.nn.Linear(200, 72).to(dev), fuse_relu=True)
    with torch.no_grad():                                    # bf16-representable operands: the fp32 oracle then
        ref.weight.copy_(ref.weight.bfloat16().float())      # sees the same ReLU mask as the bf16 kernels
    ours.load_state_dict(ref.state_dict())
    x = torch.randn(24, 200, device=dev).bfloat16().float()
    x1, x2 = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    y1 = ref(x1).relu(); y2 = ours(x2)
    assert torch.allclose(y1, y2.float(), atol=0.05, rtol=0.05)
    g = torch.randn_like(y1)
    y1.backward(g); y2.backward(g.to(y2.dtype))
    assert _rel(x2.grad, x1.grad) < 2e-2
    assert _rel(ours.weight.grad, ref.weight.grad) < 2e-2
    assert _rel(ours.bias.grad, ref.bias.grad) < 2e-2


def _ndhwc(t):
    return t.permute(0, 2, 3, 4, 1).contiguous()


@pytest.mark.parametrize('C,shape', [(16, (2, 9, 10, 13)), (32, (1, 6, 6, 6)), (128, (2, 4, 7, 5)), (256, (2, 3, 4, 3))])
def test_bn_relu_pool_block_matches_torch(dev, C, shape):
    """fused stats + BN + ReLU + MaxPool forward and backward vs the PyTorch op chain (fp32)."""
    from coinstac_dinunet_b200.ops import vbm
    torch.manual_seed(C)
    N, D, H, W = shape
    y = (torch.randn(N, D, H, W, C, device=dev) * 2 + 0.3).to(torch.bfloat16)
    gamma = torch.rand(C, device=dev) + 0.5
    gamma[::3] *= -1                                        # negative scales must work too
    beta = torch.randn(C, device=dev) * 0.1
    stats = vbm.bn_stats(y)
    yf = y.float().reshape(-1, C)
    assert torch.allclose(stats[:C], yf.sum(0), rtol=1e-3, atol=1e-2)
    rm, rv = torch.zeros(C, device=dev), torch.ones(C, device=dev)
    mean, invstd = vbm.bn_finalize(stats, yf.shape[0], 1e-5, 0.1, rm, rv)
    p = vbm.bn_relu_pool_fwd(y, mean, invstd, gamma, beta)

    y_ref = y.float().permute(0, 4, 1, 2, 3).requires_grad_(True)
    bn = torch.nn.BatchNorm3d(C).to(dev)
    with torch.no_grad():
        bn.weight.copy_(gamma); bn.bias.copy_(beta)
    p_ref = torch.nn.functional.max_pool3d(torch.relu(bn(y_ref)), 2)
    assert _rel(p, _ndhwc(p_ref)) < 1e-2
    assert torch.allclose(rm, bn.running_mean, atol=1e-3) and torch.allclose(rv, bn.running_var, rtol=1e-3, atol=1e-3)

    dp = torch.randn_like(p_ref)
    p_ref.backward(dp)
    dy, dgamma, dbeta = vbm.bn_relu_pool_bwd(y, _ndhwc(dp).to(torch.bfloat16), mean, invstd, gamma, beta)
    assert _rel(dgamma, bn.weight.grad) < 2e-2 and _rel(dbeta, bn.bias.grad) < 2e-2
    assert _rel(dy, _ndhwc(y_ref.grad)) < 2e-2
    # pass A from the pooled tensors only (xhat_argmax = (p - beta) / gamma)
    dy2, dgamma2, dbeta2 = vbm.bn_relu_pool_bwd(y, _ndhwc(dp).to(torch.bfloat16), mean, invstd, gamma, beta, p=p)
    assert _rel(dgamma2, bn.weight.grad) < 3e-2 and _rel(dbeta2, bn.bias.grad) < 2e-2
    assert _rel(dy2, _ndhwc(y_ref.grad)) < 2.5e-2


def test_native_vbmnet_forward_and_buffers_match_fp32_modules(dev):
    """whole model vs the plain fp32 PyTorch modules: logits and BatchNorm running statistics (the gradient comparison
    lives in test_hardening_gpu.py, against an oracle that quantises where the kernels do)."""
    from coinstac_dinunet_b200.models import VBMNet
    torch.manual_seed(3)
    shape = (33, 34, 35)
    ref = VBMNet(input_shape=shape).to(dev)
    nat = VBMNet(input_shape=shape, native=True).to(dev)
    nat.load_state_dict(ref.state_dict())
    assert nat.is_native
    x = torch.randn(4, 1, *shape, device=dev)
    out_ref, out_nat = ref(x), nat(x)
    assert _rel(out_nat, out_ref) < 0.08
    for (n1, b1), (_, b2) in zip(ref.named_buffers(), nat.named_buffers()):
        assert torch.allclose(b1.float(), b2.float(), rtol=5e-2, atol=5e-3), n1


@pytest.mark.parametrize('cin,cout,shape', [(16, 32, (2, 7, 9, 11)), (32, 64, (1, 6, 5, 9)), (64, 128, (2, 4, 5, 6)),
                                            (128, 256, (1, 3, 4, 3)), (32, 16, (1, 8, 8, 8)), (256, 128, (2, 3, 4, 3)),
                                            (16, 32, (3, 20, 24, 20)), (16, 32, (1, 5, 7, 60)),
                                            (64, 32, (1, 4, 9, 7))])
@pytest.mark.parametrize('impl', ['halo', 'halo2', 'tma', 'gather'])
def test_tcgen05_conv3d_matches_torch(dev, cin, cout, shape, impl, monkeypatch):
    from coinstac_dinunet_b200.ops.conv3d import conv3d_igemm_fwd, conv3d_igemm_bwd
    monkeypatch.setenv('COINN_CONV_IMPL', impl)
    torch.manual_seed(cin + cout)
    N, D, H, W = shape
    x = torch.randn(N, D, H, W, cin, device=dev).to(torch.bfloat16)
    w = (torch.randn(cout, cin, 3, 3, 3, device=dev) * (27 * cin) ** -0.5)
    y = conv3d_igemm_fwd(x, w)
    x_ref = x.float().permute(0, 4, 1, 2, 3).requires_grad_(True)
    w_ref = w.to(torch.bfloat16).float().requires_grad_(True)
    y_ref = torch.nn.functional.conv3d(x_ref, w_ref, padding=1)
    assert _rel(y, _ndhwc(y_ref)) < 1e-2
    dy = torch.randn_like(y_ref)
    y_ref.backward(dy)
    dx, dw = conv3d_igemm_bwd(_ndhwc(dy).to(torch.bfloat16), x, w, need_dx=True)
    assert _rel(dx, _ndhwc(x_ref.grad)) < 1.5e-2
    assert _rel(dw, w_ref.grad) < 1.5e-2


@pytest.mark.parametrize('cin,cout,shape', [(16, 32, (2, 7, 9, 11)), (32, 64, (1, 6, 5, 9)), (64, 128, (2, 4, 5, 6)),
                                            (128, 256, (1, 3, 4, 3)), (16, 32, (2, 20, 24, 20)), (32, 64, (2, 9, 30, 30)),
                                            (16, 32, (1, 3, 5, 60))])
@pytest.mark.parametrize('impl', ['halo', 'gather'])
def test_tcgen05_conv3d_wgrad_matches_torch(dev, cin, cout, shape, impl, monkeypatch):
    from coinstac_dinunet_b200.ops.conv3d_wgrad import conv3d_wgrad
    monkeypatch.setenv('COINN_WGRAD_IMPL', impl)
    torch.manual_seed(cin * 3 + cout)
    N, D, H, W = shape
    x = torch.randn(N, D, H, W, cin, device=dev).to(torch.bfloat16)
    dy = torch.randn(N, D, H, W, cout, device=dev).to(torch.bfloat16)
    dw = conv3d_wgrad(dy, x)
    w = torch.zeros(cout, cin, 3, 3, 3, device=dev)
    _, dw_ref, _ = torch.ops.aten.convolution_backward(
        dy.float().permute(0, 4, 1, 2, 3), x.float().permute(0, 4, 1, 2, 3), w, None, [1, 1, 1], [1, 1, 1], [1, 1, 1],
        False, [0, 0, 0], 1, [False, True, False])
    assert _rel(dw, dw_ref) < 5e-3, _rel(dw, dw_ref)


def test_cuda_graph_step_matches_eager(tmp_path):
    """whole-step CUDA graph (forward + loss + metrics + backward + fused optimizer) through the full protocol
    on one GPU: same round trace, same number of fused steps, scores written.  (Final parameters are those of the
    *selected* checkpoint; on 20-sample folds model selection can flip on a 1e-7 loss difference, so exact parameter
    equality is asserted by test_cuda_graph_replay_bit_equals_eager below instead.)"""
    import os, sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_dist_cpu import run_workers
    (tmp_path / 'eager').mkdir(); (tmp_path / 'graph').mkdir()
    eager = run_workers('protocol', tmp_path / 'eager', nproc=1, port=29721, extra=['transport=nvlink'])
    graph = run_workers('protocol', tmp_path / 'graph', nproc=1, port=29722, extra=['transport=nvlink', 'cuda_graph=1'])
    assert graph['graphed'] and not eager['graphed']
    assert graph['trace'] == eager['trace'] and graph['csv']
    assert graph['fused_steps'] == eager['fused_steps']


def test_cuda_graph_replay_bit_equals_eager(dev):
    """GraphedStep (capture with rollback, replay, ring-buffer drain) vs the eager NvlinkLearner loop."""
    from coinstac_dinunet_b200.models import FSNet
    from coinstac_dinunet_b200.models.common import ClassificationTrainer
    from coinstac_dinunet_b200.parallel.arena import DistArena
    from coinstac_dinunet_b200.parallel.graph_step import GraphedStep
    from coinstac_dinunet_b200.distrib.nodes.remote import EmptyDataHandle

    class T(ClassificationTrainer):
        def _init_nn_model(self):
            self.nn['fs_net'] = FSNet()

    def make():
        cache = {'mode': 'train', 'seed': 3, 'learning_rate': 1e-2, 'num_class': 2, 'monitor_metric': 'f1', 'gpus': [0]}
        tr = T(data_handle=EmptyDataHandle(cache, {}, {}))
        tr.init_nn(init_model=True, init_optim=True, set_devices=True, init_weights=True)
        model, opt = tr.nn['fs_net'], tr.optimizer['adam']
        arena = DistArena(model, opt, device=dev, backend='nvlink')
        learner = type('L', (), {})()
        learner.trainer, learner.arena, learner.device, learner.model = tr, arena, dev, model
        return tr, arena, learner

    g = torch.Generator().manual_seed(1)
    batches = [{'inputs': torch.randn(8, 66, generator=g), 'labels': torch.randint(0, 2, (8,), generator=g)} for _ in range(6)]
    tr1, a1, _ = make()
    tr1.nn['fs_net'].train()
    losses, tp = [], 0
    for b in batches:
        it = tr1.iteration(b)
        it['loss'].backward()
        a1.reduce_and_step()
        losses.append(float(it['loss'].detach())); tp += it['metrics'].tp + it['metrics'].tn
    tr2, a2, learner = make()
    gs = GraphedStep(learner).capture(batches[0])
    for b in batches:
        gs.step(b)
    avg, met = gs.drain()
    assert torch.equal(a1.flat_param, a2.flat_param)
    assert abs(float(avg.get()[0]) - sum(losses) / len(losses)) < 1e-4
    assert met.tp + met.tn == tp and met.tp + met.tn + met.fp + met.fn == 48
    assert int(a2.step_count) == 6 and gs.kernels_per_replay >= 3


def test_pack_conv_weights_kernel_matches_torch_packs(dev):
    """one-launch bf16 weight packing (fprop + dgrad layouts) vs the reference torch permutations."""
    from coinstac_dinunet_b200.ops import conv3d
    torch.manual_seed(0)
    for cout, cin in ((32, 16), (64, 32), (256, 128)):
        w = torch.randn(cout, cin, 3, 3, 3, device=dev)
        wf, kf, wd, kd = conv3d.pack_weights(w)
        rf, kf_ref = conv3d.pack_fprop_weight(w)
        rd, kd_ref = conv3d.pack_dgrad_weight(w)
        assert (kf, kd) == (kf_ref, kd_ref)
        assert torch.equal(wf, rf) and torch.equal(wd, rd)


def test_symm_allreduce_single_rank_is_identity(dev):
    from coinstac_dinunet_b200.parallel.arena import SymmAllReduce
    red = SymmAllReduce(1000, dev)
    ts = [torch.randn(7, 3, device=dev), torch.randn(11, device=dev)]
    want = [t.clone() for t in ts]
    red.mean_(ts)
    assert all(torch.equal(a, b) for a, b in zip(ts, want))


def test_bucketed_backward_overlap_equals_single_launch(dev):
    """per-bucket fused kernels launched from grad hooks on a side stream == one launch after backward."""
    from coinstac_dinunet_b200.parallel.arena import DistArena
    def make():
        torch.manual_seed(9)
        m = torch.nn.Sequential(torch.nn.Linear(300, 500), torch.nn.ReLU(), torch.nn.Linear(500, 400), torch.nn.ReLU(),
                                torch.nn.Linear(400, 7)).to(dev)
        return m, torch.optim.Adam(m.parameters(), lr=1e-2)
    (m1, o1), (m2, o2) = make(), make()
    a1 = DistArena(m1, o1, device=dev, backend='nvlink')
    a2 = DistArena(m2, o2, device=dev, backend='nvlink').enable_overlap(bucket_bytes=256 << 10)
    assert len(a2._overlap['buckets']) >= 2
    for step in range(5):
        x = torch.randn(32, 300, device=dev)
        m1(x).square().mean().backward(); a1.reduce_and_step()
        a2.arm_overlap(); m2(x).square().mean().backward()
        assert a2.reduce_and_step() == 'bucketed'
    torch.cuda.synchronize()
    assert torch.equal(a1.flat_param, a2.flat_param) and torch.equal(a1.m, a2.m)
    assert int(a1.step_count) == int(a2.step_count) == 5
    assert float(a2.flat_grad.abs().max()) == 0.0


def _conv1_block_reference(x, w, gamma, beta, dp=None, eps=1e-5):
    """fp32 torch oracle of the first VBM block on bf16-rounded conv operands (training-mode BatchNorm)."""
    w = w.clone().requires_grad_(True)
    gamma, beta = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    y = torch.nn.functional.conv3d(x.unsqueeze(1), w, padding=1)
    mean, var = y.mean((0, 2, 3, 4)), y.var((0, 2, 3, 4), unbiased=False)
    z = (y - mean[None, :, None, None, None]) * (var + eps).rsqrt()[None, :, None, None, None]
    z = torch.relu(z * gamma[None, :, None, None, None] + beta[None, :, None, None, None])
    p = torch.nn.functional.max_pool3d(z, 2)
    out = {'y': y.detach(), 'mean': mean.detach(), 'invstd': (var + eps).rsqrt().detach(), 'z': z.detach(), 'p': p.detach()}
    if dp is not None:
        p.backward(dp)
        out.update(dw=w.grad, dgamma=gamma.grad, dbeta=beta.grad)
    return out


C1F_SHAPES = [(2, 9, 11, 13), (1, 6, 8, 70), (1, 131, 4, 10), (2, 4, 5, 121)]


@pytest.mark.parametrize('shape', C1F_SHAPES)
def test_conv1_fused_stats_and_pool(dev, shape):
    """conv1_fused.cu STATS / POOL modes (conv recomputed, y never stored) vs the fp32 oracle."""
    from coinstac_dinunet_b200.ops import vbm
    torch.manual_seed(5)
    N, D, H, W = shape
    x = torch.randn(N, D, H, W, device=dev).bfloat16().float()
    w = (torch.randn(16, 1, 3, 3, 3, device=dev) * 0.2).bfloat16().float()
    gamma, beta = torch.rand(16, device=dev) + 0.5, torch.randn(16, device=dev) * 0.2
    ref = _conv1_block_reference(x, w, gamma, beta)
    xp = vbm.conv1_pad_input_hd(x)
    rows = xp.view(N, H + 2, D + 2, -1)
    assert torch.equal(rows[:, 1:-1, 1:-1, 1:W + 1].float(), x.permute(0, 2, 1, 3))
    assert float(rows[:, 0].abs().max()) == 0 and float(rows[..., 0].abs().max()) == 0 and float(rows[..., W + 1:].abs().max()) == 0
    stats = vbm.conv1_fused_stats(xp, w, shape)
    yf = ref['y'].permute(0, 2, 3, 4, 1).reshape(-1, 16)
    assert torch.allclose(stats[:16], yf.sum(0), rtol=1e-3, atol=5e-2)
    assert torch.allclose(stats[16:], (yf * yf).sum(0), rtol=1e-3, atol=5e-2)
    p, code = vbm.conv1_fused_pool(xp, w, ref['mean'], ref['invstd'], gamma, beta, shape)
    p_ref = ref['p'].permute(0, 2, 3, 4, 1)
    assert p.shape == p_ref.shape and code.shape == p.shape
    assert float((p.float() - p_ref).abs().max()) < 0.03 + 0.01 * float(p_ref.abs().max())
    # the code byte must point at a window element that attains the maximum, and flag ReLU activity
    PD, PH, PW = D // 2, H // 2, W // 2
    zw = ref['z'][:, :, :2 * PD, :2 * PH, :2 * PW].reshape(N, 16, PD, 2, PH, 2, PW, 2).permute(0, 2, 4, 6, 1, 3, 5, 7)
    zw = zw.reshape(N, PD, PH, PW, 16, 8)
    picked = zw.gather(-1, (code & 7).long().unsqueeze(-1)).squeeze(-1)
    assert float((picked - p_ref).abs().max()) < 2e-3
    active = (code >> 3) == 1
    assert bool(((p_ref > 1e-3) <= active).all()) and bool((active <= (p_ref > -1e-3)).all())


@pytest.mark.parametrize('shape', C1F_SHAPES)
def test_conv1_fused_block_backward(dev, shape):
    """conv1_fused.cu BWD mode: recomputed conv -> BN/ReLU/pool backward in registers -> dy in shared memory ->
    tcgen05 weight-gradient GEMM, vs autograd through the fp32 torch block."""
    from coinstac_dinunet_b200.ops import vbm
    torch.manual_seed(6)
    N, D, H, W = shape
    x = torch.randn(N, D, H, W, device=dev).bfloat16().float()
    w = (torch.randn(16, 1, 3, 3, 3, device=dev) * 0.2).bfloat16().float()
    gamma, beta = torch.rand(16, device=dev) + 0.5, torch.randn(16, device=dev) * 0.2
    dp = torch.randn(N, 16, D // 2, H // 2, W // 2, device=dev).bfloat16().float()
    ref = _conv1_block_reference(x, w, gamma, beta, dp=dp)
    rm, rv = torch.zeros(16, device=dev), torch.ones(16, device=dev)
    xin = x.clone().requires_grad_(False)
    wq, gq, bq = w.clone().requires_grad_(True), gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    p = vbm.ConvBnReluPoolFn.apply(xin, wq, gq, bq, rm, rv, 1e-5, 0.1, True, 'auto')
    assert float((p.float() - ref['p'].permute(0, 2, 3, 4, 1)).abs().max()) < 0.03 + 0.01 * float(ref['p'].abs().max())
    p.backward(dp.permute(0, 2, 3, 4, 1).contiguous().to(p.dtype))
    cos = torch.nn.functional.cosine_similarity(wq.grad.flatten(), ref['dw'].flatten(), dim=0)
    assert float(cos) > 0.999 and _rel(wq.grad, ref['dw']) < 3e-2, (float(cos), _rel(wq.grad, ref['dw']))
    assert _rel(gq.grad, ref['dgamma']) < 3e-2 and _rel(bq.grad, ref['dbeta']) < 3e-2
    assert torch.allclose(rm, 0.1 * ref['mean'], rtol=1e-3, atol=1e-4)


@pytest.mark.parametrize('cin,cout,shape', [(16, 32, (2, 6, 9, 20)), (32, 64, (1, 5, 7, 30)), (32, 16, (2, 4, 6, 11))])
def test_conv3d_halo_fused_bn_stats(dev, cin, cout, shape):
    """the halo conv epilogue also produces the BatchNorm sums of the stored output (no separate pass over y)."""
    from coinstac_dinunet_b200.ops import conv3d as c3
    torch.manual_seed(8)
    N, D, H, W = shape
    x = torch.randn(N, D, H, W, cin, device=dev).bfloat16()
    w = torch.randn(cout, cin, 3, 3, 3, device=dev) * 0.05
    y0 = c3.conv3d_igemm_fwd(x, w)
    y, stats = c3.conv3d_igemm_fwd(x, w, want_stats=True)
    assert stats is not None and torch.equal(y, y0)
    yf = y.float().reshape(-1, cout)
    assert torch.allclose(stats[:cout], yf.sum(0), rtol=1e-3, atol=1e-2)
    assert torch.allclose(stats[cout:], (yf * yf).sum(0), rtol=1e-3, atol=1e-2)


@pytest.mark.parametrize('M,K,N,relu,xdt', [(8, 9216, 256, True, torch.bfloat16), (16, 66, 256, False, torch.float32),
                                             (3, 64, 2, False, torch.float32), (32, 130, 33, True, torch.float32)])
def test_small_linear_matches_torch_and_accumulates_in_place(dev, M, K, N, relu, xdt):
    """linear_small.cu: forward/backward vs fp32 torch; dW/db accumulate directly into pre-existing .grad buffers."""
    from coinstac_dinunet_b200.ops import linear as _lin
    from coinstac_dinunet_b200.ops.linear import SmallLinearFn
    _lin.DIRECT_GRAD_DISABLED = False          # an earlier overlap test may have switched direct accumulation off
    torch.manual_seed(12)
    lin = torch.nn.Linear(K, N).to(dev)
    x0 = torch.randn(M, K, device=dev).to(xdt)
    xr = x0.float().clone().requires_grad_(True)
    yr = torch.nn.functional.linear(xr, lin.weight, lin.bias)
    yr = yr.relu() if relu else yr
    g = torch.randn_like(yr)
    gw, gb, gx = torch.autograd.grad(yr, [lin.weight, lin.bias, xr], g)
    # direct accumulation: .grad exists and already holds something
    lin.weight.grad = torch.full_like(lin.weight, 0.5)
    lin.bias.grad = torch.full_like(lin.bias, -0.25)
    wp, bp = lin.weight.grad.data_ptr(), lin.bias.grad.data_ptr()
    x1 = x0.clone().requires_grad_(True)
    y = SmallLinearFn.apply(x1, lin.weight, lin.bias, relu)
    assert torch.allclose(y, yr, rtol=2e-4, atol=2e-4)
    y.backward(g)
    assert lin.weight.grad.data_ptr() == wp and lin.bias.grad.data_ptr() == bp
    assert torch.allclose(lin.weight.grad - 0.5, gw, rtol=2e-3, atol=2e-4)
    assert torch.allclose(lin.bias.grad + 0.25, gb, rtol=2e-3, atol=2e-4)
    assert _rel(x1.grad, gx) < (1e-2 if xdt == torch.bfloat16 else 1e-4)
    # fallback: no .grad yet -> ordinary autograd accumulation
    lin.weight.grad = None; lin.bias.grad = None
    x2 = x0.clone().requires_grad_(True)
    SmallLinearFn.apply(x2, lin.weight, lin.bias, relu).backward(g)
    assert torch.allclose(lin.weight.grad, gw, rtol=2e-3, atol=2e-4) and torch.allclose(lin.bias.grad, gb, rtol=2e-3, atol=2e-4)


@pytest.mark.parametrize('first', [True, False])
def test_conv_block_direct_grad_mode_matches_autograd_mode(dev, first):
    """parameters that already own .grad buffers -> persistent accumulators + one finalize launch, None to autograd;
    must give the same gradients (added onto what .grad held) and the same running statistics as the plain mode."""
    from coinstac_dinunet_b200.ops import vbm
    from coinstac_dinunet_b200.ops import linear as _lin
    _lin.DIRECT_GRAD_DISABLED = False
    torch.manual_seed(21)
    cin, cout = (1, 16) if first else (16, 32)
    x = torch.randn(2, 8, 10, 12, device=dev) if first else torch.randn(2, 8, 10, 12, cin, device=dev).bfloat16()
    w0 = torch.randn(cout, cin, 3, 3, 3, device=dev) * 0.1
    g0, b0 = torch.rand(cout, device=dev) + 0.5, torch.randn(cout, device=dev) * 0.1
    dp = None
    out = {}
    for mode in ('plain', 'direct', 'direct'):                  # direct twice: the accumulators must come back zeroed
        w, g, b = (t.clone().requires_grad_(True) for t in (w0, g0, b0))
        rm, rv, nbt = torch.zeros(cout, device=dev), torch.ones(cout, device=dev), torch.zeros((), dtype=torch.long, device=dev)
        if mode == 'direct':
            w.grad, g.grad, b.grad = torch.full_like(w, 0.25), torch.full_like(g, 0.25), torch.full_like(b, 0.25)
        xin = x.clone().requires_grad_(not first)
        p = vbm.ConvBnReluPoolFn.apply(xin, w, g, b, rm, rv, 1e-5, 0.1, True, 'auto', nbt)
        dp = torch.randn_like(p) if dp is None else dp
        p.backward(dp)
        off = 0.25 if mode == 'direct' else 0.0
        cur = {'p': p.detach().float(), 'dw': w.grad - off, 'dg': g.grad - off, 'db': b.grad - off, 'rm': rm, 'rv': rv,
               'dx': None if first else xin.grad.float()}
        assert int(nbt) == 1
        if out:
            for k, v in cur.items():
                if v is not None:
                    assert torch.allclose(v, out[k], rtol=2e-3, atol=2e-3), (mode, k, float((v - out[k]).abs().max()))
        else:
            out = cur


@pytest.mark.parametrize('cin,cout,shape', [(64, 128, (2, 4, 5, 6)), (128, 256, (1, 3, 4, 3)), (64, 128, (2, 15, 18, 15)),
                                            (128, 256, (2, 7, 9, 7)), (64, 64, (1, 5, 20, 9))])
def test_conv3d_wgrad_tap_matches_torch(dev, cin, cout, shape, monkeypatch):
    """per-tap TMA-box weight gradient (conv3d_wgrad_tap.cu) for the C_in >= 64 blocks vs torch."""
    from coinstac_dinunet_b200.ops import conv3d_wgrad as cw
    monkeypatch.setenv('COINN_WGRAD_IMPL', 'tap')
    torch.manual_seed(cin + cout)
    N, D, H, W = shape
    x = torch.randn(N, D, H, W, cin, device=dev).to(torch.bfloat16)
    dy = torch.randn(N, D, H, W, cout, device=dev).to(torch.bfloat16)
    dw = cw.conv3d_wgrad(dy, x)
    assert cw.last_impl == 'tap'
    w = torch.zeros(cout, cin, 3, 3, 3, device=dev)
    _, dw_ref, _ = torch.ops.aten.convolution_backward(
        dy.float().permute(0, 4, 1, 2, 3), x.float().permute(0, 4, 1, 2, 3), w, None, [1, 1, 1], [1, 1, 1], [1, 1, 1],
        False, [0, 0, 0], 1, [False, True, False])
    assert _rel(dw, dw_ref) < 5e-3, _rel(dw, dw_ref)


@pytest.mark.parametrize('M,K,N,relu', [(16, 66, 256, True), (16, 256, 128, True), (8, 64, 32, True), (32, 130, 33, False),
                                        (2, 9, 5, True)])
def test_fused_linear_bn1d_relu_matches_torch(dev, M, K, N, relu):
    """Linear + BatchNorm1d (training) + ReLU in one launch each way vs the PyTorch module chain, incl. running statistics,
    num_batches_tracked, in-place .grad accumulation and eval mode."""
    from coinstac_dinunet_b200.ops.linear import linear_bn_relu
    torch.manual_seed(M + K + N)
    lin, bn = torch.nn.Linear(K, N).to(dev), torch.nn.BatchNorm1d(N).to(dev)
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5); bn.weight[::3] *= -1; bn.bias.normal_(0, 0.3)
    lin_r, bn_r = torch.nn.Linear(K, N).to(dev), torch.nn.BatchNorm1d(N).to(dev)
    lin_r.load_state_dict(lin.state_dict()); bn_r.load_state_dict(bn.state_dict())
    x = torch.randn(M, K, device=dev)
    xr = x.clone().requires_grad_(True)
    zr = bn_r(lin_r(xr))
    zr = zr.relu() if relu else zr
    g = torch.randn_like(zr)
    zr.backward(g)
    x1 = x.clone().requires_grad_(True)
    z = linear_bn_relu(x1, lin, bn, relu=relu)
    assert torch.allclose(z, zr, rtol=1e-3, atol=1e-4), float((z - zr).abs().max())
    z.backward(g)
    assert torch.allclose(bn.running_mean, bn_r.running_mean, rtol=1e-4, atol=1e-5)
    assert torch.allclose(bn.running_var, bn_r.running_var, rtol=1e-4, atol=1e-5)
    assert int(bn.num_batches_tracked) == 1
    for a, b, name in ((lin.weight.grad, lin_r.weight.grad, 'dW'), (bn.weight.grad, bn_r.weight.grad, 'dgamma'),
                       (bn.bias.grad, bn_r.bias.grad, 'dbeta'), (x1.grad, xr.grad, 'dx')):
        assert _rel(a, b) < 2e-3, (name, _rel(a, b))
    assert float(lin.bias.grad.abs().max()) < 1e-4            # a bias in front of BatchNorm has (numerically) no gradient
    # direct mode: existing fp32 .grad buffers are accumulated into in place, autograd gets None
    wp = lin.weight.grad.data_ptr()
    before = lin.weight.grad.clone()
    linear_bn_relu(x.clone(), lin, bn, relu=relu).backward(g)
    assert lin.weight.grad.data_ptr() == wp and _rel(lin.weight.grad - before, lin_r.weight.grad) < 5e-3
    # eval mode uses the running statistics (ours has seen the batch twice by now: give the reference its second pass)
    bn_r(lin_r(x))
    bn.eval(); bn_r.eval()
    with torch.no_grad():
        ze = linear_bn_relu(x, lin, bn, relu=relu)
        zer = bn_r(lin_r(x)); zer = zer.relu() if relu else zer
    assert torch.allclose(ze, zer, rtol=2e-3, atol=2e-4)


def test_native_fsnet_matches_reference_modules(dev):
    """FreeSurfer MLP on the fused Linear+BN1d+ReLU kernels vs the stock modules: logits, every gradient, BN buffers."""
    from coinstac_dinunet_b200.models import FSNet
    torch.manual_seed(2)
    ref, nat = FSNet().to(dev), FSNet(native=True).to(dev)
    nat.load_state_dict(ref.state_dict())
    ref.train(); nat.train()
    x = torch.randn(16, 66, device=dev)
    y = torch.randint(0, 2, (16,), device=dev)
    o_ref, o_nat = ref(x), nat(x)
    assert torch.allclose(o_nat, o_ref, rtol=2e-3, atol=2e-4)
    torch.nn.functional.cross_entropy(o_ref, y).backward()
    torch.nn.functional.cross_entropy(o_nat, y).backward()
    for (n1, p1), (_, p2) in zip(ref.named_parameters(), nat.named_parameters()):
        if n1.endswith('bias') and n1.startswith('features') and p1.dim() == 1 and 'features.' in n1 and int(n1.split('.')[1]) % 3 == 0:
            assert float(p2.grad.abs().max()) < 1e-4            # Linear bias under BatchNorm
            continue
        assert _rel(p2.grad, p1.grad) < 5e-3, (n1, _rel(p2.grad, p1.grad))
    for (n1, b1), (_, b2) in zip(ref.named_buffers(), nat.named_buffers()):
        assert torch.allclose(b1.float(), b2.float(), rtol=1e-4, atol=1e-5), n1


# ------------------------------------------------------------------------------------ low-rank engines (K10-K12)
@pytest.mark.parametrize('rank', [1, 2, 4])
def test_powersgd_kernels_match_reference_round(dev, rank):
    """psgd_mq / orthogonalize_batched / psgd_mtp / psgd_reconstruct over several matrices in one launch each, two
    consecutive rounds (error feedback + warm start carried), vs the PyTorch formulation of the reference's math."""
    from coinstac_dinunet_b200.ops.lowrank import PowerSGDPlan, powersgd_round_reference
    torch.manual_seed(rank)
    shapes = [(256, 66), (128, 256), (300, 1000), (2, 32), (64, 27 * 16), (17,), (256,)]       # (2, 32): min dim <= rank -> dense
    params = [torch.nn.Parameter(torch.randn(*s, device=dev)) for s in shapes]
    offsets, off = [], 0
    for p in params:
        offsets.append(off); off += (p.numel() + 7) // 8 * 8
    plan = PowerSGDPlan(params, offsets, rank, dev)
    G = torch.zeros(off, device=dev); E = torch.zeros(off, device=dev)
    P = torch.zeros(max(plan.p_numel, 4), device=dev); Q = torch.zeros(plan.q_numel + plan.low_numel + 8, device=dev)
    Qin = torch.randn(plan.q_numel, device=dev)
    mats = [(i, n, m, g, po, qo) for (i, n, m, g, po, qo) in plan.mats]
    errs = [torch.zeros(n, m, device=dev) for _, n, m, *_ in mats]
    qs = [Qin[qo:qo + m * rank].view(m, rank).clone() for _, n, m, g, po, qo in mats]
    Qcur = Qin.clone()
    for rnd in range(2):
        grads = [torch.randn(n, m, device=dev) for _, n, m, *_ in mats]
        G.zero_()
        for (_, n, m, g, *_), t in zip(mats, grads):
            G[g:g + n * m] = t.reshape(-1)
        low_vals = []
        for (i, g_off, numel) in plan.low:
            v = torch.randn(numel, device=dev); G[g_off:g_off + numel] = v; low_vals.append(v)
        approx, errs, qs = powersgd_round_reference(grads, errs, [q.clone() for q in qs], rank, lambda ts: None)
        Qbuf = torch.zeros_like(Q); Qbuf[:plan.q_numel] = Qcur
        plan.orthogonalize(Qbuf, which=1)
        plan.mq(G, E, Qbuf, P, True)
        plan.orthogonalize(P, which=0)
        plan.mtp(E, P, Q)
        plan.gather_low(G, Q)
        for (i, g_off, numel), v in zip(plan.low, low_vals):
            pass
        plan.scatter_low(Q, G)
        plan.reconstruct(G, E, P, Q, True)
        for (_, n, m, g, po, qo), a, e, q in zip(mats, approx, errs, qs):
            assert _rel(G[g:g + n * m].view(n, m), a) < 2e-4, (rnd, n, m)
            assert float((E[g:g + n * m].view(n, m) - e).norm()) < 2e-4 * float(grads[0].norm()) + 2e-4 * float(e.norm())
            assert _rel(Q[qo:qo + m * rank].view(m, rank), q) < 2e-4
        for (i, g_off, numel), v in zip(plan.low, low_vals):          # rank-1 gradients round-trip unchanged
            assert torch.equal(G[g_off:g_off + numel], v)
        Qcur = Q[:plan.q_numel].clone()


@pytest.mark.parametrize('rows_b,rows_c,n,rank', [(256, 67, 16, 10), (32, 33, 8, 10), (128, 257, 80, 10), (2, 33, 16, 10),
                                                   (64, 9217, 32, 4)])
def test_lowrank_factor_is_the_truncated_svd(dev, rows_b, rows_c, n, rank):
    """gram -> coefficient-space power iteration -> skinny GEMMs: left @ right.T is the best rank-k approximation of
    B @ C.T (compared with torch.linalg.svd), and the column-block segmented form gives the same answer."""
    from coinstac_dinunet_b200.ops.lowrank import lowrank_factor
    torch.manual_seed(n + rank)
    decay = torch.logspace(0, -3, n, device=dev)                     # a spectrum with a clear ordering
    B, C = torch.randn(rows_b, n, device=dev) * decay, torch.randn(rows_c, n, device=dev)
    left, right = lowrank_factor(B, C, rank, 40, 1e-6)
    k = left.shape[1]
    assert k == min(rank, rows_b, rows_c, n) and right.shape == (rows_c, k)
    full = B.double() @ C.double().t()
    U, S, Vh = torch.linalg.svd(full, full_matrices=False)
    best = (U[:, :k] * S[:k]) @ Vh[:k]
    got = left.double() @ right.double().t()
    opt_err = float((full - best).norm())
    assert float((full - got).norm()) <= opt_err * 1.02 + 1e-4 * float(full.norm())
    # right factors are orthonormal where kept
    gram = right.t() @ right
    keep = gram.diagonal() > 0.5
    # (exactly orthonormal only at convergence: close singular values leave O(1e-2) cross terms after 40 sweeps)
    assert torch.allclose(gram[keep][:, keep], torch.eye(int(keep.sum()), device=dev), atol=5e-2)
    if n % 4 == 0 and n >= 8:                                        # segmented operands: 4 column blocks, strided
        kseg = n // 4
        stride = (max(rows_b, rows_c) * kseg + 64)
        bufB, bufC = torch.zeros(4 * stride, device=dev), torch.zeros(4 * stride, device=dev)
        for s in range(4):
            bufB[s * stride:s * stride + rows_b * kseg] = B[:, s * kseg:(s + 1) * kseg].reshape(-1)
            bufC[s * stride:s * stride + rows_c * kseg] = C[:, s * kseg:(s + 1) * kseg].reshape(-1)
        l2, r2 = lowrank_factor(None, None, rank, 40, 1e-6, b_seg=(bufB, rows_b, n, kseg, stride), c_seg=(bufC, rows_c, n, kseg, stride))
        assert torch.allclose(l2 @ r2.t(), left @ right.t(), rtol=1e-3, atol=1e-3 * float(full.abs().max()))


@pytest.mark.parametrize('out_f,in_f,k,bias', [(256, 66, 10, True), (2, 32, 2, True), (64, 128, 16, False), (33, 9216, 10, True)])
def test_dad_reconstruct_matches_matmul(dev, out_f, in_f, k, bias):
    from coinstac_dinunet_b200.ops.lowrank import dad_reconstruct
    torch.manual_seed(k)
    delta = torch.randn(out_f, k, device=dev)
    act = torch.randn(in_f + int(bias), k, device=dev)
    wg, bg = torch.full((out_f, in_f), 7.0, device=dev), torch.full((out_f,), 7.0, device=dev)
    dad_reconstruct(delta, act, wg, bg if bias else None, scale=0.5)
    full = 0.5 * delta @ act.t()
    assert torch.allclose(wg, full[:, :in_f], rtol=1e-4, atol=1e-4)
    if bias:
        assert torch.allclose(bg, full[:, -1], rtol=1e-4, atol=1e-4)


# ------------------------------------------------------------------------------------------- MX-FP8 (config 4)
@pytest.mark.parametrize('R,K,dt', [(5, 32, torch.float32), (128, 256, torch.bfloat16), (33, 100, torch.float32), (256, 9216, torch.bfloat16)])
def test_quantize_mx_matches_reference_and_roundtrips(dev, R, K, dt):
    from coinstac_dinunet_b200.ops.fp8 import dequantize_mx, quantize_mx, quantize_mx_reference
    torch.manual_seed(R + K)
    x = (torch.randn(R, K, device=dev) * torch.logspace(-3, 3, K, device=dev)).to(dt)
    x[0, :min(K, 32)] = 0                                       # an all-zero block
    q, sf = quantize_mx(x)
    q_ref, sf_ref = quantize_mx_reference(x)
    assert torch.equal(sf, sf_ref)
    assert torch.equal(q, q_ref)
    back = dequantize_mx(q, sf, K)
    xf = x.float()
    blocks = torch.nn.functional.pad(xf, (0, q.shape[1] - K)).view(R, -1, 32)
    tol = blocks.abs().amax(-1, keepdim=True) * 2 ** -3         # e4m3: 3 mantissa bits, relative to the block max
    assert bool(((torch.nn.functional.pad(back, (0, q.shape[1] - K)).view(R, -1, 32) - blocks).abs() <= tol + 1e-30).all())


@pytest.mark.parametrize('M,N,K', [(128, 128, 128), (256, 384, 512), (100, 70, 256), (8, 256, 9216), (300, 200, 1152)])
def test_mxfp8_block_scaled_gemm_matches_dequantised_oracle(dev, M, N, K):
    """tcgen05.mma.kind::mxf8f6f4.block_scale with scale factors in TMEM vs fp32 matmul of the dequantised operands
    (exact up to fp32 accumulation order: the products of e4m3 values and power-of-two scales are exact in fp32)."""
    from coinstac_dinunet_b200.ops.fp8 import dequantize_mx, gemm_mxfp8, quantize_mx
    torch.manual_seed(M + N + K)
    a = torch.randn(M, K, device=dev) * torch.logspace(-2, 2, K, device=dev)      # scales differ from block to block
    b = torch.randn(N, K, device=dev) * torch.logspace(1, -1, K, device=dev)
    aq, asf = quantize_mx(a)
    bq, bsf = quantize_mx(b)
    ref = dequantize_mx(aq, asf).double() @ dequantize_mx(bq, bsf).double().t()
    got = gemm_mxfp8(aq, asf, bq, bsf, split_k=1)
    assert _rel(got, ref.float()) < 1e-5, _rel(got, ref.float())
    got2 = gemm_mxfp8(aq, asf, bq, bsf)                           # auto split-K
    assert _rel(got2, ref.float()) < 1e-5
    bias = torch.randn(N, device=dev)
    got3 = gemm_mxfp8(aq, asf, bq, bsf, bias=bias, relu=True, split_k=1, out_dtype=torch.bfloat16)
    assert _rel(got3, (ref.float() + bias).relu()) < 1e-2
    # and against the unquantised product: fp8 tolerance
    assert _rel(got, a @ b.t()) < 6e-2


def test_fp8_linear_autograd(dev):
    from coinstac_dinunet_b200.ops.fp8 import linear_fp8
    torch.manual_seed(0)
    lin = torch.nn.Linear(384, 200).to(dev)
    x = torch.randn(160, 384, device=dev, requires_grad=True)
    xr = x.detach().clone().requires_grad_(True)
    y = linear_fp8(x, lin.weight, lin.bias, relu=True)
    yr = torch.nn.functional.linear(xr, lin.weight, lin.bias).relu()
    assert _rel(y, yr) < 6e-2
    g = torch.randn_like(yr)
    gw_r, gb_r, gx_r = torch.autograd.grad(yr, [lin.weight, lin.bias, xr], g)
    gw, gb, gx = torch.autograd.grad(y, [lin.weight, lin.bias, x], g)
    # e4m3 carries 3 mantissa bits: ~6 % per element, products of two quantised operands on unstructured (random) data
    # average to ~10-15 % of the (cancelling) sum - direction is what matters
    for a, b in ((gw, gw_r), (gx, gx_r)):
        assert _rel(a, b) < 0.25 and torch.nn.functional.cosine_similarity(a.flatten(), b.flatten(), dim=0) > 0.97
    assert _rel(gb, gb_r) < 0.25          # ReLU mask decided by the fp8 forward: a few % of the gates differ


def _dequant_grouped(q, sf, box):
    R, K = q.shape
    G = box // 32
    e = sf.view(torch.uint8).view(R, K // box, 4)[:, :, :G].reshape(R, K // 32).float()
    return (q.view(torch.float8_e4m3fn).float().view(R, K // 32, 32) * torch.exp2(e - 127).unsqueeze(-1)).view(R, K)


@pytest.mark.parametrize('cin,cout,shape', [(32, 64, (2, 6, 7, 9)), (64, 128, (1, 5, 6, 20)), (128, 256, (2, 3, 4, 5)),
                                            (256, 128, (1, 3, 4, 3)), (64, 32, (1, 4, 9, 7)), (32, 64, (2, 30, 36, 30))])
def test_mxfp8_conv3d_matches_dequantised_oracle(dev, cin, cout, shape):
    """Block-scaled fp8 implicit-GEMM conv (per-tap TMA boxes, scale factors gathered at the shifted voxels and staged in
    TMEM) vs torch conv3d on the dequantised operands: only fp32 accumulation order and the bf16 output rounding differ."""
    from coinstac_dinunet_b200.ops import conv3d as c3
    torch.manual_seed(cin + cout)
    N, D, H, W = shape
    x = (torch.randn(N, D, H, W, cin, device=dev) * torch.logspace(-1, 1, cin, device=dev)).bfloat16()
    w = torch.randn(cout, cin, 3, 3, 3, device=dev) * (27 * cin) ** -0.5
    y = c3.conv3d_igemm_fwd(x, w, fp8=True)
    assert c3.last_impl == 'mxfp8'
    box = min(cin, 128)
    xq, xsf = c3._quantize_rows(x.reshape(-1, cin), cin, box)
    wk = w.permute(0, 2, 3, 4, 1).reshape(cout, 27 * cin).bfloat16()
    wq, wsf = c3._quantize_rows(wk, 27 * cin, box)
    xd = _dequant_grouped(xq, xsf, box).view(N, D, H, W, cin).permute(0, 4, 1, 2, 3)
    wd = _dequant_grouped(wq, wsf, box).view(cout, 3, 3, 3, cin).permute(0, 4, 1, 2, 3)
    ref = torch.nn.functional.conv3d(xd.double(), wd.double(), padding=1).float()
    assert _rel(y, _ndhwc(ref)) < 6e-3, _rel(y, _ndhwc(ref))
    # against the unquantised conv: fp8 tolerance
    ref_hp = torch.nn.functional.conv3d(x.float().permute(0, 4, 1, 2, 3), w.bfloat16().float(), padding=1)
    assert _rel(y, _ndhwc(ref_hp)) < 6e-2


def test_fp8_conv_block_trains_like_bf16(dev):
    """config 4 numerics: a 32->64 block with fp8 fprop/dgrad (bf16 wgrad, fp32 master weights) follows the bf16 block."""
    from coinstac_dinunet_b200.ops import vbm
    torch.manual_seed(12)
    x = torch.randn(2, 8, 10, 12, 32, device=dev).bfloat16()
    w0 = torch.randn(64, 32, 3, 3, 3, device=dev) * 0.05
    g0, b0 = torch.rand(64, device=dev) + 0.5, torch.randn(64, device=dev) * 0.1
    outs = {}
    for backend in ('auto', 'fp8'):
        w, g, b = (t.clone().requires_grad_(True) for t in (w0, g0, b0))
        rm, rv = torch.zeros(64, device=dev), torch.ones(64, device=dev)
        xin = x.clone().requires_grad_(True)
        p = vbm.ConvBnReluPoolFn.apply(xin, w, g, b, rm, rv, 1e-5, 0.1, True, backend)
        dp = torch.ones_like(p) * 0.01 + (torch.arange(p.numel(), device=dev).view_as(p) % 7).to(p.dtype) * 0.01
        p.backward(dp)
        outs[backend] = (p.float(), w.grad.clone(), g.grad.clone(), xin.grad.float())
    for a, b_, name in zip(outs['fp8'], outs['auto'], ('p', 'dw', 'dgamma', 'dx')):
        cos = torch.nn.functional.cosine_similarity(a.flatten(), b_.flatten(), dim=0)
        assert _rel(a, b_) < (0.12 if name == 'p' else 0.6) and cos > 0.85, (name, _rel(a, b_), float(cos))
