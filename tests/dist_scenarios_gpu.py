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
        a2.enable_overlap(bucket_bytes=128 << 10)
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
        mine = a2.flat_param.clone()
        allp = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allp, mine)
        results.append({'variant': variant, 'how': how, 'err': err, 'buckets': len(a2._overlap['buckets']),
                        'identical': all(torch.equal(allp[0], q) for q in allp[1:]),
                        'steps': int(a2.step_count), 'zeroed': float(a2.flat_grad.abs().max()) == 0.0})
    if rank == 0:
        with open(os.path.join(work, 'result.json'), 'w') as fp:
            json.dump({'results': results, 'world': world}, fp)


SCENARIOS = {'fused': scenario_fused, 'allreduce': scenario_allreduce, 'overlap': scenario_overlap}
