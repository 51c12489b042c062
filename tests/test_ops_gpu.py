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
