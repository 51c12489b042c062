"""Correctness evidence asked for by the round-1 review: kernels at the benchmark shapes, the whole native model against
an oracle that quantises at the same points, and convergence of the native model on separable data."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def dev():
    from coinstac_dinunet_b200 import ops
    assert ops.native_available()
    return torch.device('cuda', 0)


def _rel(a, b):
    a, b = a.float(), b.float()
    return float((a - b).norm() / (b.norm() + 1e-12))


def _ndhwc(t):
    return t.permute(0, 2, 3, 4, 1).contiguous()


# ------------------------------------------------------------------------- bf16-emulating oracle
class _Round(torch.autograd.Function):
    """value stored as bf16 on the way forward AND its gradient stored as bf16 on the way back."""

    @staticmethod
    def forward(ctx, x):
        return x.bfloat16().float()

    @staticmethod
    def backward(ctx, g):
        return g.bfloat16().float()


class _RoundGrad(torch.autograd.Function):
    """forward untouched (value never stored), gradient tile written as bf16 (first block: dy lives in shared memory)."""

    @staticmethod
    def forward(ctx, x):
        return x.clone()

    @staticmethod
    def backward(ctx, g):
        return g.bfloat16().float()


def _bf16_weight(w):
    return w + (w.bfloat16().float() - w).detach()       # forward sees the bf16 operand, gradient stays fp32


def emulated_forward(model, x):
    """fp32 PyTorch forward of VBMNet with the quantisation points of the native path: bf16 conv operands, stored conv
    outputs (blocks 2-5) / pooled outputs / their gradients in bf16, BatchNorm statistics over the stored values, fp32
    head fed with the bf16 pooled features."""
    F = torch.nn.functional
    h = _Round.apply(x)
    for i, blk in enumerate(model.blocks):
        y = F.conv3d(h, _bf16_weight(blk.conv.weight), padding=1)
        y = _RoundGrad.apply(y) if i == 0 else _Round.apply(y)
        y = F.batch_norm(y, None, None, blk.bn.weight, blk.bn.bias, True, 0.0, blk.bn.eps)
        h = _Round.apply(F.max_pool3d(torch.relu(y), 2))
    z = h.flatten(1)
    for layer in model.head:
        z = layer(z)
    return model.classifier(z)


def _grads(model, x, y, fwd=None):
    model.zero_grad(set_to_none=True)
    out = fwd(model, x) if fwd else model(x)
    torch.nn.functional.cross_entropy(out.float(), y).backward()
    return out.detach().float(), {n: p.grad.detach().clone() for n, p in model.named_parameters()}


def test_native_vbmnet_matches_bf16_emulated_oracle(dev):
    """The round-1 whole-model test accepted cos > 0.9 / rel < 0.45 against an fp32 oracle.  Here the oracle rounds where
    the kernels round, and the comparison is calibrated:

    * the well-conditioned quantities - logits and the classifier gradients - agree to 2e-2 (measured 2e-3);
    * the conv-stack gradients of a randomly initialised BN network are chaotic in the rounding noise (measured on B200,
      profiles/r2/determinism.txt: a 1e-7 perturbation of the oracle itself moves the first block's gradient by 1e-3; the
      1e-6 run-to-run jitter of fp32-atomic BatchNorm sums moves the native one by several percent although every kernel
      is bit-reproducible on identical inputs).  For those the bound is 3x the native run-to-run spread + 2e-2: the
      kernels must sit inside their own noise envelope around the oracle."""
    from coinstac_dinunet_b200.models import VBMNet
    torch.manual_seed(3)
    shape = (33, 34, 35)
    ref = VBMNet(input_shape=shape).to(dev)
    nat = VBMNet(input_shape=shape, native=True).to(dev)
    nat.load_state_dict(ref.state_dict())
    ref.train(); nat.train()
    y = torch.randint(0, 2, (4,), device=dev)
    x = torch.randn(4, 1, *shape, device=dev) + 0.5 * (y.float() * 2 - 1).view(-1, 1, 1, 1, 1)
    state = {k: v.clone() for k, v in nat.state_dict().items()}
    out_nat, g_nat = _grads(nat, x, y)
    nat.load_state_dict(state)
    out_nat2, g_nat2 = _grads(nat, x, y)
    out_ref, g_ref = _grads(ref, x, y, emulated_forward)
    assert _rel(out_nat, out_ref) < 2e-2, _rel(out_nat, out_ref)
    report = {}
    for n in g_ref:
        assert torch.isfinite(g_nat[n]).all(), n
        err, jitter = _rel(g_nat[n], g_ref[n]), _rel(g_nat2[n], g_nat[n])
        tight = n.startswith('classifier')                     # measured 1.5e-3; everything upstream inherits the flip noise
        bound = 2e-2 if tight else 3 * jitter + 2e-2
        report[n] = (round(err, 4), round(jitter, 4), round(bound, 4))
        cos = torch.nn.functional.cosine_similarity(g_nat[n].flatten(), g_ref[n].flatten(), dim=0)
        assert err < bound and cos > 0.95, (n, report[n], float(cos))


def test_native_vbmnet_converges_on_separable_data(dev):
    """Training with the native kernels + fused optimizer actually learns: class-dependent intensity shift (the kind of
    signal a small-receptive-field CNN with pooling can see), 40 Adam steps; the stock fp32 model on the same stream of
    batches is the yardstick."""
    from coinstac_dinunet_b200 import ops
    from coinstac_dinunet_b200.models import VBMNet
    from coinstac_dinunet_b200.parallel.arena import DistArena
    shape = (33, 34, 35)

    def train(native):
        torch.manual_seed(0)
        model = VBMNet(input_shape=shape, native=native).to(dev)
        opt = torch.optim.Adam(model.parameters(), lr=1e-3)
        arena = DistArena(model, opt, device=dev, backend='nvlink') if native else None
        g = torch.Generator(device='cpu').manual_seed(1)
        model.train()
        losses, accs = [], []
        for step in range(40):
            yb = torch.randint(0, 2, (8,), generator=g).to(dev)
            xb = torch.randn(8, 1, *shape, generator=g).to(dev) + (yb.float() * 2 - 1).view(-1, 1, 1, 1, 1) * 0.5
            if native:
                loss, pred = ops.softmax_nll(model(xb), yb)
                loss.backward()
                arena.reduce_and_step()
            else:
                out = model(xb)
                loss, pred = torch.nn.functional.cross_entropy(out, yb), out.argmax(1)
                loss.backward()
                opt.step(); opt.zero_grad()
            losses.append(float(loss)); accs.append(float((pred == yb).float().mean()))
        return losses, accs

    ref_l, ref_a = train(False)
    assert sum(ref_l[-5:]) / 5 < 0.3, ('the yardstick itself did not learn', ref_l[-5:])
    nat_l, nat_a = train(True)
    assert sum(nat_l[-5:]) / 5 < 0.35 < sum(nat_l[:3]) / 3, (nat_l[:3], nat_l[-5:], ref_l[-5:])
    assert sum(nat_a[-5:]) / 5 >= 0.9, nat_a[-5:]


# ----------------------------------------------------------------------- benchmark-shape kernels
def test_block2_convs_at_benchmark_shape(dev):
    """8 x 60x72x60 x 16 -> 32 (the layer that is 45 % of the step): fprop / dgrad / wgrad vs cuDNN on the same bf16
    operands; exercises tile tails, the persistent scheduler and >2^31-byte-free indexing at the real size."""
    from coinstac_dinunet_b200.ops.conv3d import conv3d_igemm_fwd, conv3d_igemm_bwd
    torch.manual_seed(1)
    N, D, H, W, cin, cout = 8, 60, 72, 60, 16, 32
    x = torch.randn(N, D, H, W, cin, device=dev).bfloat16()
    w = torch.randn(cout, cin, 3, 3, 3, device=dev) * (27 * cin) ** -0.5
    y, stats = conv3d_igemm_fwd(x, w, want_stats=True)
    xr = x.permute(0, 4, 1, 2, 3)                                     # channels_last_3d view, bf16
    wr = w.bfloat16().contiguous(memory_format=torch.channels_last_3d)
    y_ref = torch.nn.functional.conv3d(xr, wr, padding=1).float()
    assert _rel(y, _ndhwc(y_ref)) < 1e-2
    if stats is not None:
        yf = y.float().reshape(-1, cout)
        assert torch.allclose(stats[:cout], yf.sum(0), rtol=2e-3, atol=1.0)
        assert torch.allclose(stats[cout:], (yf * yf).sum(0), rtol=2e-3, atol=1.0)
    dy = torch.randn(N, D, H, W, cout, device=dev).bfloat16()
    dx, dw = conv3d_igemm_bwd(dy, x, w, need_dx=True)
    dx_ref, dw_ref, _ = torch.ops.aten.convolution_backward(
        dy.permute(0, 4, 1, 2, 3), xr, wr, None, [1, 1, 1], [1, 1, 1], [1, 1, 1], False, [0, 0, 0], 1, [True, True, False])
    assert _rel(dx, _ndhwc(dx_ref.float())) < 1.5e-2
    assert _rel(dw, dw_ref.float()) < 1.5e-2


def test_first_block_at_benchmark_shape(dev):
    """8 x 121x145x121 through the fused first block (statistics / BN+ReLU+pool / backward, conv recomputed, nothing stored
    at full resolution) vs the fp32 torch block on bf16-rounded operands."""
    from coinstac_dinunet_b200.ops import vbm
    torch.manual_seed(2)
    N, D, H, W = 8, 121, 145, 121
    x = torch.randn(N, D, H, W, device=dev).bfloat16().float()
    w = (torch.randn(16, 1, 3, 3, 3, device=dev) * 0.2).bfloat16().float()
    gamma, beta = torch.rand(16, device=dev) + 0.5, torch.randn(16, device=dev) * 0.2
    dp = torch.randn(N, 16, D // 2, H // 2, W // 2, device=dev).bfloat16().float()
    wr, gr, br = w.clone().requires_grad_(True), gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    yr = torch.nn.functional.conv3d(x.unsqueeze(1), wr, padding=1)
    zr = torch.nn.functional.batch_norm(yr, None, None, gr, br, True, 0.0, 1e-5)
    pr = torch.nn.functional.max_pool3d(torch.relu(zr), 2)
    pr.backward(dp)
    mean_ref = yr.detach().mean((0, 2, 3, 4))
    del yr, zr
    rm, rv = torch.zeros(16, device=dev), torch.ones(16, device=dev)
    wq, gq, bq = w.clone().requires_grad_(True), gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    p = vbm.ConvBnReluPoolFn.apply(x, wq, gq, bq, rm, rv, 1e-5, 0.1, True, 'auto')
    p_ref = pr.detach().permute(0, 2, 3, 4, 1)
    assert float((p.float() - p_ref).abs().max()) < 0.03 + 0.01 * float(p_ref.abs().max())
    p.backward(dp.permute(0, 2, 3, 4, 1).contiguous().to(p.dtype))
    assert _rel(wq.grad, wr.grad) < 3e-2, _rel(wq.grad, wr.grad)
    assert _rel(gq.grad, gr.grad) < 3e-2 and _rel(bq.grad, br.grad) < 3e-2
    assert torch.allclose(rm, 0.1 * mean_ref, rtol=1e-3, atol=1e-4)


def test_whole_step_at_benchmark_shape_is_finite_and_stable(dev):
    """One full native step at 8 x 1x121x145x121 twice from the same state.  Every kernel is bit-reproducible on identical
    inputs (scripts/diag_stats.py); the fp32-atomic BatchNorm sums jitter at 1e-6 and the bf16 rounding flips they cause
    reach the logits at < 1e-2 and the well-conditioned (classifier) gradient at < 1e-2."""
    from coinstac_dinunet_b200 import ops
    from coinstac_dinunet_b200.models import VBMNet
    torch.manual_seed(4)
    model = VBMNet(native=True).to(dev)
    model.train()
    y = torch.randint(0, 2, (8,), device=dev)
    x = torch.randn(8, 1, 121, 145, 121, device=dev) + 0.5 * (y.float() * 2 - 1).view(-1, 1, 1, 1, 1)
    state = {k: v.clone() for k, v in model.state_dict().items()}
    runs = []
    for _ in range(2):
        model.load_state_dict(state)
        model.zero_grad(set_to_none=True)
        out = model(x)
        loss, _ = ops.softmax_nll(out, y)
        loss.backward()
        assert torch.isfinite(loss)
        runs.append((out.detach().float().clone(), {n: p.grad.clone() for n, p in model.named_parameters()}))
    for n, g in runs[0][1].items():
        assert torch.isfinite(g).all(), n
    assert _rel(runs[1][0], runs[0][0]) < 2e-2
    assert _rel(runs[1][1]['classifier.weight'], runs[0][1]['classifier.weight']) < 1e-2
    # every gradient keeps its direction between the two runs
    for n in runs[0][1]:
        a, b = runs[0][1][n].flatten(), runs[1][1][n].flatten()
        assert torch.nn.functional.cosine_similarity(a, b, dim=0) > 0.97, n


# ------------------------------------------------------------------------------ user-defined models
def _user_net():
    from torch import nn
    return nn.Sequential(
        nn.Conv3d(1, 8, 3, padding=1), nn.BatchNorm3d(8), nn.ReLU(), nn.MaxPool3d(2),
        nn.Conv3d(8, 24, 3, padding=1, bias=False), nn.BatchNorm3d(24), nn.ReLU(), nn.MaxPool3d(2),
        nn.Conv3d(24, 48, 3, padding=1, bias=False), nn.BatchNorm3d(48), nn.ReLU(), nn.MaxPool3d(2),
        nn.Flatten(), nn.Linear(48 * 4 * 4 * 4, 64), nn.BatchNorm1d(64), nn.ReLU(), nn.Linear(64, 2))


def test_nativize_user_cnn_matches_torch_forward(dev):
    """A user-defined 3-block CNN (channel counts the kernels are NOT instantiated for: 1->8->24->48) routed through
    ops.nativize: same logits as the stock modules within bf16 tolerance, same BatchNorm buffers, gradients for every
    parameter."""
    from coinstac_dinunet_b200.ops import nativize
    torch.manual_seed(11)
    ref, nat = _user_net().to(dev), _user_net().to(dev)
    nat.load_state_dict(ref.state_dict())
    report = []
    nativize(nat, report=report)
    assert any(r[1] == 'conv_stack' for r in report) and any(r[1] == 'linear_bn_relu' for r in report)
    ref.train(); nat.train()
    y = torch.randint(0, 2, (8,), device=dev)
    x = torch.randn(8, 1, 32, 32, 32, device=dev) + 0.5 * (y.float() * 2 - 1).view(-1, 1, 1, 1, 1)
    from coinstac_dinunet_b200 import ops
    l0 = ops.launch_count
    o_ref, o_nat = ref(x), nat(x)
    assert ops.launch_count - l0 >= 8                      # conv blocks + fused linear layers really ran natively
    assert _rel(o_nat, o_ref) < 8e-2, _rel(o_nat, o_ref)
    torch.nn.functional.cross_entropy(o_nat, y).backward()
    for n, p in nat.named_parameters():
        if n == '0.bias':
            continue          # a Conv3d bias in front of BatchNorm has no effect and no gradient: the native block ignores it
        assert p.grad is not None and torch.isfinite(p.grad).all(), n
    for (n1, b1), (_, b2) in zip(ref.named_buffers(), nat.named_buffers()):
        assert torch.allclose(b1.float(), b2.float(), rtol=5e-2, atol=5e-3), n1
    nat.eval(); ref.eval()
    with torch.no_grad():
        assert _rel(nat(x), ref(x)) < 8e-2


def test_nativize_user_cnn_trains(dev):
    """... and it trains: the nativized user model + DistArena fused optimizer learns the separable task."""
    from coinstac_dinunet_b200.ops import nativize
    from coinstac_dinunet_b200.parallel.arena import DistArena
    torch.manual_seed(5)
    model = nativize(_user_net().to(dev))
    opt = torch.optim.Adam(model.parameters(), lr=2e-3)
    arena = DistArena(model, opt, device=dev, backend='nvlink')
    g = torch.Generator(device='cpu').manual_seed(3)
    model.train()
    losses = []
    for step in range(40):
        yb = torch.randint(0, 2, (8,), generator=g).to(dev)
        xb = torch.randn(8, 1, 32, 32, 32, generator=g).to(dev) + (yb.float() * 2 - 1).view(-1, 1, 1, 1, 1) * 0.5
        loss = torch.nn.functional.cross_entropy(model(xb), yb)
        loss.backward()
        arena.reduce_and_step()
        losses.append(float(loss))
    assert sum(losses[-5:]) / 5 < 0.1 and losses[0] > 5 * (sum(losses[-5:]) / 5), (losses[:3], losses[-5:])


def test_bucketed_overlap_with_direct_grad_kernels(dev):
    """The native blocks write gradients straight into the arena and hand autograd None, so bucket launches are driven by
    ops.linear.notify_grad_written instead of autograd hooks.
    (a) identical gradients through both launch schedules -> bit-identical parameters and moments;
    (b) the real thing: buckets launched on the side stream while backward is still running, eager and captured."""
    from coinstac_dinunet_b200 import ops
    from coinstac_dinunet_b200.models import VBMNet
    from coinstac_dinunet_b200.ops.linear import notify_grad_written
    from coinstac_dinunet_b200.parallel.arena import DistArena
    shape = (33, 34, 35)

    def make(overlap):
        torch.manual_seed(0)
        m = VBMNet(input_shape=shape, native=True).to(dev)
        a = DistArena(m, torch.optim.Adam(m.parameters(), lr=1e-3), device=dev, backend='nvlink')
        if overlap:
            a.enable_overlap(bucket_bytes=1 << 20)
            assert len(a._overlap['buckets']) >= 2
        m.train()
        return m, a

    (m1, a1), (m2, a2) = make(False), make(True)
    g = torch.Generator(device='cpu').manual_seed(5)
    # (a) same gradient values, two schedules
    for step in range(3):
        grads = torch.randn(a1.numel, generator=g).to(dev)
        a1.flat_grad.copy_(grads); a2.flat_grad.copy_(grads)
        a1.reduce_and_step()
        a2.arm_overlap()
        for p in reversed(a2.params):                      # "backward" order: every parameter reports its gradient final
            notify_grad_written(p)
        assert a2.reduce_and_step() == 'bucketed'
    torch.cuda.synchronize()
    assert torch.equal(a1.flat_param, a2.flat_param) and torch.equal(a1.m, a2.m) and torch.equal(a1.v, a2.v)
    assert int(a1.step_count) == int(a2.step_count) == 3 and float(a2.flat_grad.abs().max()) == 0.0
    # (b) real backward passes with the buckets in flight
    start = a2.flat_param.clone()
    for step in range(3):
        y = torch.randint(0, 2, (4,), generator=g).to(dev)
        x = torch.randn(4, 1, *shape, generator=g).to(dev) + 0.5 * (y.float() * 2 - 1).view(-1, 1, 1, 1, 1).to(dev)
        for m, a in ((m1, a1), (m2, a2)):
            loss, _ = ops.softmax_nll(m(x), y)
            a.arm_overlap()
            loss.backward()
            how = a.reduce_and_step()
        assert how == 'bucketed'
    torch.cuda.synchronize()
    u1, u2 = (a1.flat_param - start).flatten(), (a2.flat_param - start).flatten()
    assert torch.nn.functional.cosine_similarity(u1, u2, dim=0) > 0.7        # same trajectory up to bf16-flip noise
    dirty = {n: float(p.grad.abs().max()) for n, p in m2.named_parameters() if float(p.grad.abs().max()) != 0.0}
    assert int(a2.step_count) == 6 and not dirty, dirty
    # and captured: the bucket launch on the side stream is recorded as a parallel branch of the graph
    x = torch.randn(4, 1, *shape, device=dev); y = torch.randint(0, 2, (4,), device=dev)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        loss, _ = ops.softmax_nll(m2(x), y); a2.arm_overlap(); loss.backward(); a2.reduce_and_step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    import gc
    del loss
    gc.collect()
    graph = torch.cuda.CUDAGraph()
    before = a2.flat_param.clone()
    with torch.cuda.graph(graph):
        loss, _ = ops.softmax_nll(m2(x), y); a2.arm_overlap(); loss.backward(); a2.reduce_and_step()
    graph.replay()
    torch.cuda.synchronize()
    assert not torch.equal(before, a2.flat_param) and float(a2.flat_grad.abs().max()) == 0.0


def test_fp8_training_tracks_bf16_over_200_steps(dev):
    """BASELINE config 4 numerics: VBMNet with MX-FP8 block-scaled conv fprop/dgrad (blocks 3-5) vs the bf16 kernels on the
    same stream of batches, 200 fused-Adam steps: both learn the task and the smoothed loss curves stay together."""
    from coinstac_dinunet_b200 import ops
    from coinstac_dinunet_b200.models import VBMNet
    from coinstac_dinunet_b200.ops import conv3d as c3
    from coinstac_dinunet_b200.parallel.arena import DistArena
    shape = (33, 34, 35)

    def train(backend):
        torch.manual_seed(0)
        model = VBMNet(input_shape=shape, native=True, conv_backend=backend).to(dev)
        arena = DistArena(model, torch.optim.Adam(model.parameters(), lr=5e-4), device=dev, backend='nvlink')
        g = torch.Generator(device='cpu').manual_seed(1)
        model.train()
        losses, used = [], set()
        for step in range(200):
            yb = torch.randint(0, 2, (8,), generator=g).to(dev)
            xb = torch.randn(8, 1, *shape, generator=g).to(dev) + (yb.float() * 2 - 1).view(-1, 1, 1, 1, 1) * 0.3
            loss, _ = ops.softmax_nll(model(xb), yb)
            used.add(c3.last_impl)
            loss.backward()
            arena.reduce_and_step()
            losses.append(float(loss.detach()))
        return torch.tensor(losses), used

    bf16, _ = train('auto')
    fp8, used = train('fp8')
    assert 'mxfp8' in used
    smooth = lambda t: t.view(20, 10).mean(1)
    a, b = smooth(bf16), smooth(fp8)
    assert float(a[-3:].mean()) < 0.3 and float(b[-3:].mean()) < 0.3, (a.tolist(), b.tolist())
    assert float((a - b).abs().max()) < 0.3, (a.tolist(), b.tolist())
