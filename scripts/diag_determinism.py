"""Run-to-run and oracle deviations of the native VBMNet per parameter (diagnostic, not a test)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import torch
from coinstac_dinunet_b200 import ops
from coinstac_dinunet_b200.models import VBMNet
from test_hardening_gpu import emulated_forward, _rel

dev = torch.device('cuda', 0)


def run(model, x, y, fwd=None):
    model.zero_grad(set_to_none=True)
    out = fwd(model, x) if fwd else model(x)
    loss = torch.nn.functional.cross_entropy(out.float(), y)
    loss.backward()
    return out.detach().float(), {n: p.grad.detach().clone() for n, p in model.named_parameters()}


for shape, batch in (((33, 34, 35), 4), ((121, 145, 121), 8)):
    for signal in (0.0, 0.5):
        torch.manual_seed(3)
        ref = VBMNet(input_shape=shape).to(dev)
        nat = VBMNet(input_shape=shape, native=True).to(dev)
        nat.load_state_dict(ref.state_dict())
        ref.train(); nat.train()
        y = torch.randint(0, 2, (batch,), device=dev)
        x = torch.randn(batch, 1, *shape, device=dev) + signal * (y.float() * 2 - 1).view(-1, 1, 1, 1, 1)
        state = {k: v.clone() for k, v in nat.state_dict().items()}
        o1, g1 = run(nat, x, y)
        nat.load_state_dict(state)
        o2, g2 = run(nat, x, y)
        print(f'--- shape {shape} batch {batch} signal {signal}: logits run-to-run {_rel(o2, o1):.2e}')
        big = shape[0] > 100
        if not big:
            oe, ge = run(ref, x, y, emulated_forward)
            ref.load_state_dict({k: v for k, v in state.items()})
            # second oracle: same quantisation points, conv computed by a different algorithm (channels_last_3d)
            ref2 = VBMNet(input_shape=shape).to(dev).to(memory_format=torch.channels_last_3d)
            ref2.load_state_dict(state); ref2.train()
            oe2, ge2 = run(ref2, x.contiguous(memory_format=torch.channels_last_3d), y, emulated_forward)
            print(f'    logits native-vs-oracle {_rel(o1, oe):.2e}  oracle-vs-oracle2 {_rel(oe2, oe):.2e}')
        for n in g1:
            line = f'    {n:28s} run-to-run {_rel(g2[n], g1[n]):.2e}'
            if not big:
                line += f'  native-vs-oracle {_rel(g1[n], ge[n]):.2e}  oracle-vs-oracle2 {_rel(ge2[n], ge[n]):.2e}'
            print(line)
