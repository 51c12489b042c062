import sys, torch
sys.path.insert(0, '.')
from coinstac_dinunet_b200.models import FSNet
from coinstac_dinunet_b200.parallel.arena import DistArena
from coinstac_dinunet_b200 import ops
dev = torch.device('cuda')

def make():
    torch.manual_seed(0)
    m = FSNet().to(dev)
    o = torch.optim.Adam(m.parameters(), lr=1e-2)
    return m, o, DistArena(m, o, device=dev, backend='nvlink')

g = torch.Generator().manual_seed(1)
batches = [(torch.randn(8, 66, generator=g).to(dev), torch.randint(0, 2, (8,), generator=g).to(dev)) for _ in range(6)]

def step(m, a, x, y):
    loss, pred = ops.softmax_nll(m(x), y)
    loss.backward()
    a.reduce_and_step()
    return loss

# eager
m1, o1, a1 = make(); m1.train()
le = [float(step(m1, a1, x, y).detach()) for x, y in batches]
# graph
m2, o2, a2 = make(); m2.train()
sx, sy = batches[0][0].clone(), batches[0][1].clone()
snap = [t.clone() for t in (a2.flat_param, a2.m, a2.v, a2.step_count)]
bufs = [(b, b.clone()) for b in m2.buffers()]
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for _ in range(2): step(m2, a2, sx, sy)
torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
gr = torch.cuda.CUDAGraph()
with torch.cuda.graph(gr):
    gl = step(m2, a2, sx, sy)
for d, s in zip((a2.flat_param, a2.m, a2.v, a2.step_count), snap): d.copy_(s)
for b, s in bufs: b.copy_(s)
a2.flat_grad.zero_()
lg = []
for x, y in batches:
    sx.copy_(x); sy.copy_(y); gr.replay(); lg.append(float(gl.detach()))
print('eager', [round(v, 5) for v in le]); print('graph', [round(v, 5) for v in lg])
print('param diff', float((a1.flat_param - a2.flat_param).abs().max()), 'steps', int(a1.step_count), int(a2.step_count))
print('grad views intact', all(p.grad.data_ptr() == a2.flat_grad.data_ptr() + off * 4 for p, off in zip(a2.params, a2.offsets)))
