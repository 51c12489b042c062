#!/usr/bin/env python
"""Gradient-reduce sweep (BASELINE.json config 5): fused reduce+Adam kernels vs NCCL, 1 KB - 1 GB.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 bench_reduce.py [--max-mb 1024]

For every bucket size B (bytes of fp32 gradients per site) and every variant it times, on the device
(CUDA events, max over ranks), ONE full dSGD data-plane step:
    one_shot / two_shot / nvls   coinn::fused_reduce_opt_kernel  (reduce + 1/S + Adam + re-zero, no NCCL)
    nccl+fused                    torch.distributed.all_reduce(NCCL) + the S==1 fused Adam kernel   (R1)
    nccl_only                     torch.distributed.all_reduce alone                                 (R2)
and reports achieved algorithmic bandwidth B/t plus the fraction of the roofline of BASELINE.md §4:
    t_min = max(link_bytes / 770 GB/s (measured peer copy), hbm_bytes / 6578.7 GB/s (MEASURED_PEAKS))
    link_bytes: one-shot (S-1)*B | two-shot 2*(S-1)/S*B | nvls (1 + (S-1)/S)*B/... see code ; hbm ~ 28 B/param.
One JSON line per (size, variant) on rank 0; the table is also written to gpurun_out/reduce_sweep_n<S>.json.
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LINK_GBS = 770.0          # measured peer-copy bandwidth per direction (B200_PROFILING.md)
HBM_GBS = 6578.7          # MEASURED_PEAKS.json hbm_gbs


def roofline_us(nbytes, world, variant):
    nparam = nbytes / 4
    hbm = 28.0 * nparam / world if variant in ('two_shot', 'nvls') else 28.0 * nparam
    hbm += nbytes                                     # the local gradient read / re-zero
    if world == 1:
        link = 0.0
    elif variant == 'one_shot':
        link = (world - 1) * nbytes
    elif variant == 'two_shot':
        link = 2.0 * (world - 1) / world * nbytes
    else:                                              # nvls: reduced shard in, full parameters out (multicast)
        link = nbytes / world + nbytes * (world - 1) / world
    return max(link / (LINK_GBS * 1e3), hbm / (HBM_GBS * 1e3))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--max-mb', type=float, default=1024)
    ap.add_argument('--min-kb', type=float, default=1)
    ap.add_argument('--iters', type=int, default=10)
    ap.add_argument('--burst', type=int, default=10, help='back-to-back launches per timed burst (removes host launch skew: '
                                                           'after the first launch the ranks are aligned by the kernel barrier)')
    ap.add_argument('--wire16', type=int, default=1, help='also sweep the bf16-wire variants')
    a = ap.parse_args()
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1'); os.environ.setdefault('MASTER_PORT', '29590')
    os.environ.setdefault('RANK', '0'); os.environ.setdefault('WORLD_SIZE', '1'); os.environ.setdefault('LOCAL_RANK', '0')
    local = int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = torch.device('cuda', local)
    from coinstac_dinunet_b200.parallel.arena import DistArena

    sizes, b = [], int(a.min_kb * 1024)
    while b <= a.max_mb * 1024 * 1024:
        sizes.append(b)
        b *= 4
    rows = []
    for nbytes in sizes:
        n = nbytes // 4
        variants = ['one_shot', 'two_shot', 'nvls', 'nccl+fused', 'nccl_only'] if world > 1 else ['one_shot']
        if world > 1 and a.wire16:
            variants += ['two_shot@bf16', 'nvls@bf16']
        for variant in variants:
            if variant == 'one_shot' and nbytes * world > (1 << 31):
                continue                               # (S-1)*B ingress: pointless beyond the latency regime
            model = torch.nn.ParameterList([torch.nn.Parameter(torch.zeros(n, device=dev))])
            opt = torch.optim.Adam(model.parameters(), lr=1e-3)
            backend = 'nccl' if variant.startswith('nccl') else 'nvlink'
            kernel_variant, _, wire = variant.partition('@')
            arena = DistArena(model, opt, device=dev, backend=backend,
                              variant=kernel_variant if backend == 'nvlink' else 'auto', grad_dtype=wire or 'f32')
            used = variant
            iters = a.iters if nbytes <= (64 << 20) else max(3, a.iters // 3)
            burst = a.burst if nbytes <= (64 << 20) else max(2, a.burst // 3)

            def once():
                if variant == 'nccl_only':
                    dist.all_reduce(arena.flat_grad)
                    return 'nccl_only'
                return arena.reduce_and_step()
            for _ in range(3):
                arena.flat_grad.normal_()
                used = once()
            times = []
            for _ in range(iters):
                arena.flat_grad.normal_()              # fresh gradients (also evicts nothing: B up to 1 GB >> L2)
                dist.barrier(device_ids=[local])
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _b in range(burst):                # back to back on the stream, like the steps of a CUDA graph
                    once()
                e1.record()
                torch.cuda.synchronize()
                t = torch.tensor([e0.elapsed_time(e1) * 1e3 / burst], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                times.append(float(t))
            times.sort()
            med = times[len(times) // 2]
            wire_bytes = nbytes // 2 if wire else nbytes
            roof = roofline_us(nbytes, world, used if used in ('one_shot', 'two_shot', 'nvls') else 'two_shot')
            if wire:                                   # half the link bytes on the reduce leg, parameters still fp32
                roof = max(roof * 0.75, 28.0 * (nbytes / 4) / world / (HBM_GBS * 1e3))
            row = {'bytes': nbytes, 'sites': world, 'variant': variant, 'used': used, 'us_median': round(med, 2),
                   'us_min': round(times[0], 2), 'algbw_GBps': round(nbytes / med / 1e3, 2),
                   'roofline_us': round(roof, 2), 'frac_of_roofline': round(roof / med, 4)}
            rows.append(row)
            if rank == 0:
                print(json.dumps(row), flush=True)
            del arena, opt, model
            torch.cuda.empty_cache()
    if rank == 0:
        os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
        with open(os.path.join(ROOT, 'gpurun_out', f'reduce_sweep_n{world}.json'), 'w') as fp:
            json.dump(rows, fp, indent=1)
    dist.barrier(device_ids=[local])
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
