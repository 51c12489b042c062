"""GPU scenarios for tests/dist_worker.py (multi-GPU, launched through torch.distributed.run)."""
import json
import os
import time

import torch
import torch.distributed as dist


def scenario_fused(work, opts):
    """fused_reduce_opt.cu across S GPUs vs torch.optim.Adam on the exactly averaged gradient;
    all variants, odd sizes, several steps, random per-rank delays, replica bit-equality."""
    from coinstac_dinunet_b200.parallel.arena import DistArena
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = torch.device('cuda', torch.cuda.current_device())
    results = []
    for variant in ('one_shot', 'two_shot', 'nvls'):
        for width in (3, 64, 1000, 2049):
            torch.manual_seed(42)                                   # identical init on all ranks
            mk = lambda: torch.nn.Sequential(torch.nn.Linear(width, 17), torch.nn.Tanh(), torch.nn.Linear(17, width)).to(dev)
            ours, ref = mk(), mk()
            ref.load_state_dict(ours.state_dict())
            o_opt, r_opt = torch.optim.Adam(ours.parameters(), lr=1e-2), torch.optim.Adam(ref.parameters(), lr=1e-2)
            arena = DistArena(ours, o_opt, device=dev, backend='nvlink', variant=variant)
            used = None
            for step in range(4):
                g = torch.Generator(device='cpu').manual_seed(1000 * step + rank)
                x = torch.randn(8, width, generator=g).to(dev)
                ours(x).square().mean().backward()
                ref(x).square().mean().backward()
                flat = torch.cat([p.grad.reshape(-1) for p in ref.parameters()])
                parts = [torch.empty_like(flat) for _ in range(world)]
                dist.all_gather(parts, flat)
                mean = torch.stack(parts).sum(0) / world            # fixed rank order, like the kernel
                off = 0
                for p in ref.parameters():
                    p.grad.copy_(mean[off:off + p.numel()].view_as(p)); off += p.numel()
                r_opt.step(); r_opt.zero_grad()
                if (step + rank) % 2:
                    time.sleep(0.01 * rank)                         # skew the ranks: barriers must hold
                used = arena.reduce_and_step()
            torch.cuda.synchronize()
            err = max(float((a - b).abs().max()) for a, b in zip(ours.parameters(), ref.parameters()))
            mine = arena.flat_param.clone()
            allp = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(allp, mine)
            identical = all(torch.equal(allp[0], q) for q in allp[1:])
            zeroed = float(arena.flat_grad.abs().max()) == 0.0
            results.append({'variant': variant, 'used': used, 'width': width, 'err': err,
                            'identical': identical, 'zeroed': zeroed})
    if rank == 0:
        with open(os.path.join(work, 'result.json'), 'w') as fp:
            json.dump({'results': results, 'world': world}, fp)


def scenario_allreduce(work, opts):
    """SymmAllReduce (fused kernel with opt_kind NONE) vs torch.distributed.all_reduce, all variants."""
    from coinstac_dinunet_b200.parallel.arena import SymmAllReduce
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = torch.device('cuda', torch.cuda.current_device())
    results = []
    for variant in ('one_shot', 'two_shot', 'nvls'):
        red = SymmAllReduce(1 << 18, dev, variant=variant)
        for shapes in ([(5,), (3, 7)], [(1000, 33)], [(64, 1), (1,), (17, 17, 3)]):
            g = torch.Generator(device='cpu').manual_seed(rank * 100 + len(shapes))
            ts = [torch.randn(*s, generator=g).to(dev) for s in shapes]
            want = [t.clone() for t in ts]
            for t in want:
                dist.all_reduce(t)
                t /= world
            for rep in range(3):                                   # repeated use: buffers are re-zeroed by the kernel
                got = [t.clone() for t in ts]
                red.mean_(got)
            err = max(float((a - b).abs().max()) for a, b in zip(got, want))
            results.append({'variant': variant, 'err': err})
    if rank == 0:
        with open(os.path.join(work, 'result.json'), 'w') as fp:
            json.dump({'results': results, 'world': world}, fp)


def scenario_overlap(work, opts):
    """bucketed launch from grad hooks (side stream, during backward) vs one launch after backward."""
    from coinstac_dinunet_b200.parallel.arena import DistArena
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = torch.device('cuda', torch.cuda.current_device())
    results = []
    for variant in ('one_shot', 'two_shot', 'auto'):
        torch.manual_seed(5)
        mk = lambda: torch.nn.Sequential(torch.nn.Linear(257, 600), torch.nn.ReLU(), torch.nn.Linear(600, 300),
                                         torch.nn.ReLU(), torch.nn.Linear(300, 5)).to(dev)
        m1, m2 = mk(), mk()
        m2.load_state_dict(m1.state_dict())
        a1 = DistArena(m1, torch.optim.Adam(m1.parameters(), lr=1e-2), device=dev, backend='nvlink', variant=variant)
        a2 = DistArena(m2, torch.optim.Adam(m2.parameters(), lr=1e-2), device=dev, backend='nvlink', variant=variant)
        a2.enable_overlap(bucket_bytes=64 << 10)
        for step in range(6):
            g = torch.Generator(device='cpu').manual_seed(77 * step + rank)
            x = torch.randn(16, 257, generator=g).to(dev)
            m1(x).square().mean().backward()
            a1.reduce_and_step()
            if (step + rank) % 2:
                time.sleep(0.005 * (rank + 1))
            a2.arm_overlap()
            m2(x).square().mean().backward()
            how = a2.reduce_and_step()
        torch.cuda.synchronize()
        err = float((a1.flat_param - a2.flat_param).abs().max())
        # checkpoint path: sharded moments gathered per launch unit must agree between the two launch schedules
        a1.gather_state(); a2.gather_state()
        torch.cuda.synchronize()
        state_err = max(float((a1.m - a2.m).abs().max()), float((a1.v - a2.v).abs().max()))
        mine = a2.flat_param.clone()
        allp = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allp, mine)
        results.append({'variant': variant, 'how': how, 'err': err, 'state_err': state_err, 'buckets': len(a2._overlap['buckets']),
                        'identical': all(torch.equal(allp[0], q) for q in allp[1:]),
                        'steps': int(a2.step_count), 'zeroed': float(a2.flat_grad.abs().max()) == 0.0})
    if rank == 0:
        with open(os.path.join(work, 'result.json'), 'w') as fp:
            json.dump({'results': results, 'world': world}, fp)

class _Flat(torch.nn.Module):
    """One flat parameter of `n` elements: the fused kernel sees exactly the requested size."""

    def __init__(self, n, dev):
        super().__init__()
        self.w = torch.nn.Parameter(torch.zeros(n, device=dev))


def _adam_ref(p, m, v, g, t, lr=1e-2, b1=0.9, b2=0.999, eps=1e-8):
    m.mul_(b1).add_(g, alpha=1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    denom = (v.sqrt() / (1 - b2 ** t) ** 0.5).add_(eps)
    p.addcdiv_(m, denom, value=-lr / (1 - b1 ** t))


def scenario_sizes(work, opts):
    """fused reduce + Adam on 1 KB, 1 MB + 4 B, 64 MB (and 1 GB with big=1) of gradients, every variant, two steps,
    against a plain torch Adam on the exactly averaged gradient."""
    from coinstac_dinunet_b200.parallel.arena import DistArena
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = torch.device('cuda', torch.cuda.current_device())
    sizes = [256, (1 << 18) + 1, 1 << 24] + ([1 << 28] if opts.get('big') == '1' else [])
    results = []
    for n in sizes:
        for variant in ('one_shot', 'two_shot', 'nvls'):
            if variant == 'one_shot' and n > (1 << 24):
                continue
            model = _Flat(n, dev)
            with torch.no_grad():
                model.w.copy_(torch.linspace(-1, 1, n, device=dev))
            arena = DistArena(model, torch.optim.Adam(model.parameters(), lr=1e-2), device=dev, backend='nvlink', variant=variant)
            p, m, v = model.w.detach().clone(), torch.zeros(n, device=dev), torch.zeros(n, device=dev)
            used = None
            for t in (1, 2):
                gen = torch.Generator(device=dev).manual_seed(1000 * t + rank)
                g = torch.randn(n, device=dev, generator=gen) * (rank + 1)
                model.w.grad.copy_(g)
                mean = g.clone()
                dist.all_reduce(mean)
                mean /= world
                _adam_ref(p, m, v, mean, t)
                del g, mean
                used = arena.reduce_and_step()
            torch.cuda.synchronize()
            err = float((model.w.detach() - p).abs().max())
            zeroed = float(arena.flat_grad.abs().max()) == 0.0
            mine = arena.flat_param[:n].clone()
            ref0 = mine.clone()
            dist.broadcast(ref0, src=0)
            same = torch.tensor([float(torch.equal(ref0, mine))], device=dev)
            dist.all_reduce(same, op=dist.ReduceOp.MIN)
            arena.check_health()
            results.append({'numel': n, 'variant': variant, 'used': used, 'err': err, 'zeroed': zeroed,
                            'identical': bool(same.item())})
            del arena, model, p, m, v, mine, ref0
            torch.cuda.empty_cache()
    if rank == 0:
        with open(os.path.join(work, 'result.json'), 'w') as fp:
            json.dump({'results': results, 'world': world}, fp)


def scenario_wire16(work, opts):
    """precision_bits = 16 on the NVLink transport: gradients cross the wire as fp16 / bf16 (packed inside the fused
    kernel), are summed in fp32 and the optimizer runs on fp32 masters.  Oracle: Adam on the mean of the ROUNDED
    per-site gradients (what the reference computes: learner.py:17 casts, reducer.py:29 means)."""
    from coinstac_dinunet_b200.parallel.arena import DistArena
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = torch.device('cuda', torch.cuda.current_device())
    results = []
    for wire, tdt in (('bf16', torch.bfloat16), ('f16', torch.float16)):
        for variant in ('one_shot', 'two_shot', 'nvls'):
            for n in (1000, (1 << 20) + 4):
                model = _Flat(n, dev)
                arena = DistArena(model, torch.optim.Adam(model.parameters(), lr=1e-2), device=dev, backend='nvlink',
                                  variant=variant, grad_dtype=wire)
                assert arena.wire_buf is not None and arena.grad_code in (1, 2)
                p, m, v = model.w.detach().clone(), torch.zeros(n, device=dev), torch.zeros(n, device=dev)
                for t in (1, 2, 3):
                    gen = torch.Generator(device=dev).manual_seed(77 * t + rank)
                    g = torch.randn(n, device=dev, generator=gen)
                    model.w.grad.copy_(g)
                    mean = g.to(tdt).float()
                    dist.all_reduce(mean)
                    mean /= world
                    _adam_ref(p, m, v, mean, t)
                    used = arena.reduce_and_step()
                torch.cuda.synchronize()
                err = float((model.w.detach() - p).abs().max())
                zeroed = float(arena.flat_grad.abs().max()) == 0.0
                mine = arena.flat_param[:n].clone()
                ref0 = mine.clone()
                dist.broadcast(ref0, src=0)
                same = torch.tensor([float(torch.equal(ref0, mine))], device=dev)
                dist.all_reduce(same, op=dist.ReduceOp.MIN)
                results.append({'wire': wire, 'variant': variant, 'used': used, 'numel': n, 'err': err, 'zeroed': zeroed,
                                'identical': bool(same.item())})
    if rank == 0:
        with open(os.path.join(work, 'result.json'), 'w') as fp:
            json.dump({'results': results, 'world': world}, fp)


def scenario_watchdog(work, opts):
    """A site that never arrives must not hang the box: rank `world-1` launches its step 3 s late, the others run with a
    0.5 s barrier timeout.  Their kernels give up, record which peer was missing and `check_health` raises; the late
    rank then finds everybody's flags already posted and terminates too."""
    from coinstac_dinunet_b200.parallel.arena import DistArena
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = torch.device('cuda', torch.cuda.current_device())
    model = _Flat(4096, dev)
    arena = DistArena(model, torch.optim.Adam(model.parameters(), lr=1e-2), device=dev, backend='nvlink',
                      variant='one_shot', timeout_ms=500)
    model.w.grad.fill_(1.0)
    arena.reduce_and_step()                      # a healthy step first
    torch.cuda.synchronize()
    arena.check_health()
    dist.barrier()
    late = world - 1
    t0 = time.time()
    if rank == late:
        time.sleep(3.0)
    model.w.grad.fill_(1.0)
    arena.reduce_and_step()
    torch.cuda.synchronize()
    waited = time.time() - t0
    raised, msg = False, ''
    try:
        arena.check_health()
    except RuntimeError as exc:
        raised, msg = True, str(exc)
    out = [None] * world
    dist.all_gather_object(out, {'rank': rank, 'raised': raised, 'msg': msg, 'waited': waited})
    if rank == 0:
        with open(os.path.join(work, 'result.json'), 'w') as fp:
            json.dump({'results': out, 'world': world, 'late': late}, fp)


SCENARIOS = {'fused': scenario_fused, 'allreduce': scenario_allreduce, 'overlap': scenario_overlap,
             'sizes': scenario_sizes, 'wire16': scenario_wire16, 'watchdog': scenario_watchdog}
