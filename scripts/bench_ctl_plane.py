#!/usr/bin/env python
"""Round-trip cost of the DistEngine control plane (one gather + one broadcast of a small dict per round), CPU only:

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 scripts/bench_ctl_plane.py

Prints the median / p99 per round for the shared-memory mailbox (engine/shm_plane.py) and for the gloo object collectives it
replaces on one node."""
import os
import sys
import time

import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def rounds(gather, broadcast, rank, n):
    msg = {'phase': 'computation', 'mode': 'train', 'reduce': True, 'site': rank, 'scores': [0.5] * 8}
    out = []
    for i in range(n):
        t0 = time.perf_counter()
        got = gather(msg)
        broadcast(({'phase': 'computation', 'update': True, 'n': len(got) if got else 0}, False) if rank == 0 else None)
        out.append(time.perf_counter() - t0)
    out.sort()
    return out[len(out) // 2] * 1e6, out[int(len(out) * 0.99)] * 1e6


def main():
    dist.init_process_group('gloo')
    rank, world = dist.get_rank(), dist.get_world_size()
    from coinstac_dinunet_b200.engine.shm_plane import ShmMailbox
    box = [None]
    mb = ShmMailbox(None, 0, world, create=True) if rank == 0 else None
    box[0] = mb.name if mb else None
    dist.broadcast_object_list(box, src=0)
    if rank:
        mb = ShmMailbox(box[0], rank, world)
    dist.barrier()
    rounds(mb.gather, mb.broadcast, rank, 200)
    shm = rounds(mb.gather, mb.broadcast, rank, 2000)

    def g(obj):
        got = [None] * world if rank == 0 else None
        dist.gather_object(obj, got, dst=0)
        return got

    def b(obj):
        bx = [obj]
        dist.broadcast_object_list(bx, src=0)
        return bx[0]
    rounds(g, b, rank, 50)
    gloo = rounds(g, b, rank, 500)
    if rank == 0:
        print(f'ranks {world}: shared-memory mailbox median {shm[0]:.1f} us p99 {shm[1]:.1f} us | '
              f'gloo object collectives median {gloo[0]:.1f} us p99 {gloo[1]:.1f} us')
    dist.barrier()
    mb.close()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
