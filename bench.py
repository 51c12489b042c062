#!/usr/bin/env python
"""Headline benchmark (BASELINE.json): samples/sec, whole box, VBM-3D-CNN dSGD, 1/2/4/8 B200.

    python bench.py --gpus N --steps K --warmup W            # our engine
    python bench.py --impl reference --gpus N ...             # unmodified reference (baseline/_ref)

N > 1 is launched by ``python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N``:
one rank per GPU == one federated site per GPU.  Both arms train the same architecture
(coinstac_dinunet_b200.models.VBMNet == baseline/ref_models.RefVBMNet) with the same per-site batch
on synthetic volumes of the named shape (1x121x145x121) with random-init weights.

Our arm goes through the public API the whole way: ``DistEngine`` drives ``COINNLocal`` /
``COINNRemote`` rounds; with ``transport='nvlink'`` one round is ``steps_per_round`` fused steps
(forward/backward + fused cross-GPU reduce + Adam in one kernel).  Two measurements:
  * ``value``  - K steps, device-timed (CUDA events, max over ranks), inputs resident on the device
                 (4+ rotating batches, working set >> L2);
  * ``e2e``    - the same K steps where every step copies its batch from pinned host memory
                 (H2D) and reads the step's loss back (D2H).
The line printed by rank 0 follows the driver's contract.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

VBM_SHAPE = (1, 121, 145, 121)
FS_SHAPE = (66,)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=10)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--model', default='vbm', choices=['vbm', 'fs'])
    ap.add_argument('--batch', type=int, default=None, help='per-site batch (default 8 vbm / 16 fs)')
    ap.add_argument('--dtype', default='bf16', choices=['bf16', 'fp32'])
    ap.add_argument('--transport', default='nvlink', choices=['nvlink', 'nccl'])
    ap.add_argument('--variant', default='auto', choices=['auto', 'one_shot', 'two_shot', 'nvls'])
    ap.add_argument('--overlap', type=int, default=0, help='bucketed reduce+update launched from grad hooks during backward')
    ap.add_argument('--native', type=int, default=1, help='use the hand-written sm_100a model kernels')
    ap.add_argument('--skip-e2e', action='store_true')
    ap.add_argument('--input-dtype', default='bf16', choices=['bf16', 'fp32'],
                    help='dtype of the volumes in pinned host memory (bf16 halves the PCIe bytes; compute is bf16 anyway)')
    ap.add_argument('--graph', type=int, default=1, help='capture the whole step in a CUDA graph')
    return ap.parse_args()


# ----------------------------------------------------------------------------------- clocks
class ClockSampler:
    """`nvidia-smi` clocks + throttle reasons sampled every 25 ms during the timed region."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix='.csv')
            os.close(fd)
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-i', str(self.index), '-lms', '25'],
                                         stdout=open(self.path, 'w'), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        if self.proc is None:
            return out
        try:
            self.proc.terminate()
            self.proc.wait(timeout=5)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(',')]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, val in zip(names, f[5:9]):
                    if val.lower().startswith('active'):
                        reasons.add(name)
            os.remove(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), samples=len(sm))
        out['reasons'] = sorted(reasons)
        return out


def dist_setup(n):
    import torch
    import torch.distributed as dist
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    os.environ.setdefault('MASTER_PORT', '29577')
    os.environ.setdefault('RANK', '0')
    os.environ.setdefault('LOCAL_RANK', '0')
    os.environ.setdefault('WORLD_SIZE', '1')
    local = int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    if not dist.is_initialized():
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    assert dist.get_world_size() == n, f'--gpus {n} but WORLD_SIZE={dist.get_world_size()} (launch with torchrun)'
    return dist.get_rank(), local


def timed(fn, dist, torch):
    """barrier + sync | events around fn | sync + barrier; returns max-over-ranks milliseconds."""
    dist.barrier(device_ids=[torch.cuda.current_device()])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    result = fn()
    e1.record()
    torch.cuda.synchronize()
    dist.barrier(device_ids=[torch.cuda.current_device()])
    ms = torch.tensor([e0.elapsed_time(e1)], device='cuda')
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item()), result


# ------------------------------------------------------------------------------------- ours
def run_ours(a):
    import torch
    import torch.distributed as dist
    rank, local = dist_setup(a.gpus)
    from coinstac_dinunet_b200 import ops
    from coinstac_dinunet_b200.config.keys import Phase
    from coinstac_dinunet_b200.engine import DistEngine
    from coinstac_dinunet_b200.models import FSVTrainer, InMemorySynthetic, VBMTrainer
    from coinstac_dinunet_b200.models.common import pinned_collate

    shape = VBM_SHAPE if a.model == 'vbm' else FS_SHAPE
    batch = a.batch or (8 if a.model == 'vbm' else 16)
    trainer_cls = VBMTrainer if a.model == 'vbm' else FSVTrainer
    n_distinct = 4 * batch                                   # 4 rotating batches: 272 MB fp32 for VBM (> L2)
    n_files = batch * (a.warmup + a.steps + 4) * 2

    class Volumes(InMemorySynthetic):
        def __init__(self, **kw):
            super().__init__(shape=shape, num_class=2, seed=100 + rank, pin=True,
                             dtype=torch.bfloat16 if (a.input_dtype == 'bf16' and a.model == 'vbm') else torch.float32, **kw)

    work = tempfile.mkdtemp(prefix='coinn_bench_') if rank == 0 else None
    box = [work]
    dist.broadcast_object_list(box, src=0)
    work = box[0]
    spec = dict(task_id='bench', mode='train', data_dir='data', num_class=2, batch_size=batch,
                split_ratio=[0.98, 0.01, 0.01], epochs=10 ** 6, gpus=[local], agg_engine='dSGD', seed=11,
                monitor_metric='f1', learning_rate=1e-3, validation_epochs=10 ** 9, transport=a.transport, reduce_variant=a.variant,
                compute_dtype=a.dtype, channels_last='3d' if a.model == 'vbm' else None, native_ops=bool(a.native),
                input_shape=list(shape), input_size=shape[0], synthetic_distinct=n_distinct,
                reference_order=True, pin_memory=False, collate_fn=pinned_collate, cuda_graph=bool(a.graph))
    eng = DistEngine(work, inputspec=spec)
    data_dir = os.path.join(eng.state['baseDirectory'], 'data')
    os.makedirs(data_dir, exist_ok=True)
    for i in range(n_files):
        open(os.path.join(data_dir, f's{rank:02d}_{i:06d}'), 'w').close()

    from coinstac_dinunet_b200 import COINNLocal, COINNRemote, COINNDataHandle

    def local_fn(site, cache, inp, state):
        return COINNLocal(cache=cache, input=inp, state=state)(None, trainer_cls, Volumes, COINNDataHandle)

    def remote_fn(cache, inp, state):
        return COINNRemote(cache=cache, input=inp, state=state)(None, trainer_cls)

    # init_runs -> next_run: the model is built, the first fused round (1 step) runs
    eng.cache['steps_per_round'] = 1
    for _ in range(4):
        eng.step(local_fn, remote_fn)
        if '_arena' in eng.cache:
            break
    assert '_arena' in eng.cache, 'engine never reached the training phase'

    def rounds(k, resident, readback):
        eng.cache['steps_per_round'] = k
        eng.cache['readback_per_step'] = readback
        eng.cache['synthetic_device'] = f'cuda:{local}' if resident else None
        # end-to-end mode: stage batch t+1 on a copy stream while step t computes (public dataloader option)
        eng.cache['prefetch_to_device'] = None if resident else f'cuda:{local}'
        ds = eng.cache['dataset'].get('train')
        if ds is not None and (ds._x is None or (ds._x.is_cuda != resident)):
            ds._x = None                                     # re-materialise on the requested side
        eng.cache['cursor'] = 0                               # fresh iterator over the (re)placed data
        # validation_epochs is huge in the spec, so the aggregator answers every finished round
        # ("epoch") with mode=train and the next round trains again
        eng.step(local_fn, remote_fn)

    sampler = ClockSampler(local)

    # ---- device-resident inputs: the kernel-side number ----
    rounds(a.warmup, True, False)
    l0 = ops.launch_count
    sampler.start()
    ms, _ = timed(lambda: rounds(a.steps, True, False), dist, torch)
    clocks = sampler.stop()
    launches = ops.launch_count - l0

    # ---- end to end: pinned host -> device every step, loss read back every step ----
    e2e = None
    if not a.skip_e2e:
        rounds(max(a.warmup, 3), False, True)
        ms_e2e, _ = timed(lambda: rounds(a.steps, False, True), dist, torch)
        item = 2 if (a.input_dtype == 'bf16' and a.model == 'vbm') else 4
        h2d = batch * (int(torch.tensor(shape).prod()) * item + 8)
        e2e = {'value': a.gpus * batch * a.steps / (ms_e2e / 1e3), 'unit': 'samples/s',
               'ms_per_step': ms_e2e / a.steps, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': 4}

    arena = eng.cache['_arena']
    if rank == 0:
        value = a.gpus * batch * a.steps / (ms / 1e3)
        line = {
            'metric': 'samples/sec (whole box, max over sites) VBM-3D-CNN dSGD' if a.model == 'vbm'
            else 'samples/sec (whole box, max over sites) FreeSurfer-MLP dSGD',
            'value': value, 'unit': 'samples/s', 'n_gpus': a.gpus, 'steps': a.steps, 'warmup': a.warmup,
            'ms_per_step': ms / a.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': a.dtype, 'data': 'synthetic', 'impl': 'ours',
            'config': {'model': 'VBMNet 5x[Conv3d-BN-ReLU-MaxPool] + 3 FC, 3.55M params' if a.model == 'vbm'
                       else 'FSNet MLP 66-256-128-64-32-2',
                       'input': list(shape), 'per_site_batch': batch, 'global_batch': batch * a.gpus,
                       'parallelism': f'dSGD sites={a.gpus} (1 site/GPU)', 'optimizer': 'Adam(1e-3)',
                       'transport': arena.backend, 'reduce_variant': arena._pick_variant(arena.numel * 4), 'overlap_backward': bool(a.overlap),
                       'native_model_kernels': bool(a.native), 'cuda_graph': bool(a.graph),
                       'host_input_dtype': a.input_dtype if a.model == 'vbm' else 'fp32',
                       'l2_policy': 'inputs rotate over 4 distinct batches; activations per step >> 126 MB L2'},
            'clocks': clocks, 'e2e': e2e, 'gpu_launches': launches,
        }
        emit(line)
    dist.barrier(device_ids=[local])
    dist.destroy_process_group()


# ------------------------------------------------------------------------------------ stdout
_JSON_FD = None


def _guard_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner on the
    first communicator), so everything that is not the result goes to stderr: fd 1 is pointed at fd 2 for the run
    and the JSON line is written to the saved descriptor."""
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)


def emit(obj):
    data = (json.dumps(obj) + '\n').encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


# -------------------------------------------------------------------------------- reference
def run_reference(a):
    try:
        sys.path.insert(0, os.path.join(ROOT, 'baseline'))
        import ref_runner
        ref_runner.import_reference()
    except Exception as exc:  # noqa: BLE001
        emit({'impl': 'reference', 'unavailable': f'{type(exc).__name__}: {exc}'[:300]})
        return
    import torch
    import torch.distributed as dist
    rank, local = dist_setup(a.gpus)
    shape = VBM_SHAPE if a.model == 'vbm' else FS_SHAPE
    batch = a.batch or (8 if a.model == 'vbm' else 16)
    work = tempfile.mkdtemp(prefix='coinn_ref_') if rank == 0 else None
    box = [work]
    dist.broadcast_object_list(box, src=0)
    n_files = int(batch * (a.warmup + a.steps + 8) / 0.98) + 64
    eng = ref_runner.RefEngine(box[0], a.model, shape, batch, n_files, use_gpu=True)
    eng.advance_to_training()
    for _ in range(a.warmup):
        eng.round()
    sampler = ClockSampler(local)
    sampler.start()

    def k_rounds():
        for _ in range(a.steps):
            eng.round()
    ms, _ = timed(k_rounds, dist, torch)
    clocks = sampler.stop()
    if rank == 0:
        value = a.gpus * batch * a.steps / (ms / 1e3)
        h2d = batch * (int(torch.tensor(shape).prod()) * 4 + 8)
        nparam = sum(p.numel() for p in eng.cache['nn']['model'].parameters())
        emit({
            'metric': 'samples/sec (whole box, max over sites) VBM-3D-CNN dSGD' if a.model == 'vbm'
            else 'samples/sec (whole box, max over sites) FreeSurfer-MLP dSGD',
            'value': value, 'unit': 'samples/s', 'n_gpus': a.gpus, 'steps': a.steps, 'warmup': a.warmup,
            'ms_per_step': ms / a.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'fp32', 'data': 'synthetic', 'impl': 'reference',
            'config': {'model': 'RefVBMNet' if a.model == 'vbm' else 'RefFSNet', 'input': list(shape),
                       'per_site_batch': batch, 'global_batch': batch * a.gpus,
                       'parallelism': f'dSGD sites={a.gpus} (file + JSON transport, stock COINNLearner/COINNReducer)',
                       'params': nparam},
            'clocks': clocks,
            'e2e': {'value': value, 'unit': 'samples/s', 'h2d_bytes_per_step': h2d + nparam * 4,
                    'd2h_bytes_per_step': nparam * 4 + 4,
                    'note': 'the reference has no device-only mode: every step is host<->device + files'},
            'gpu_launches': 0})
    dist.barrier(device_ids=[local])
    dist.destroy_process_group()


if __name__ == '__main__':
    _guard_stdout()
    args = parse()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)
