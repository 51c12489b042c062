#!/usr/bin/env python
"""Headline benchmark (BASELINE.json): samples/sec, whole box, VBM-3D-CNN dSGD, 1/2/4/8 B200.

    python bench.py --gpus N --steps K --warmup W            # our engine
    python bench.py --impl reference --gpus N ...             # unmodified reference (baseline/_ref)
    python bench.py --impl nccl_cudnn --gpus N ...            # R1: same nn.Module on cuDNN/cuBLAS bf16 + NCCL + fused Adam
    python bench.py --model fs ...                            # BASELINE config 2 (FreeSurfer MLP) instead of the VBM CNN

N > 1 is launched by ``python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N``:
one rank per GPU == one federated site per GPU.  Both arms train the same architecture
(coinstac_dinunet_b200.models.VBMNet == baseline/ref_models.RefVBMNet) with the same per-site batch
on synthetic volumes of the named shape (1x121x145x121) with random-init weights.

Our arm goes through the public API the whole way: ``DistEngine`` drives ``COINNLocal`` /
``COINNRemote`` rounds; with ``transport='nvlink'`` one round is ``steps_per_round`` fused steps
(forward/backward + fused cross-GPU reduce + Adam in one kernel).  Two measurements:
  * ``value``  - K steps, device-timed (CUDA events, max over ranks), inputs resident on the device
                 (4+ rotating batches, working set >> L2);
  * ``e2e``    - the same K steps where every step copies its batch (fp32 volumes, the dtype the
                 reference arm moves) from pinned host memory (H2D) and reads the step's loss back
                 (D2H); ``e2e_bf16_host`` repeats it with the volumes stored as bf16 on the host.
``vs_baseline`` = value / the R1 number measured on this hardware for the same N
(``baseline/r1_measured.json``, written from ``--impl nccl_cudnn`` runs).
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
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference', 'nccl_cudnn'])
    ap.add_argument('--model', default='vbm', choices=['vbm', 'fs'])
    ap.add_argument('--batch', type=int, default=None, help='per-site batch (default 8 vbm / 16 fs)')
    ap.add_argument('--dtype', default='bf16', choices=['bf16', 'fp32', 'fp8'],
                    help="fp8 = BASELINE config 4: MX-FP8 block-scaled conv fprop/dgrad (blocks 3-5), bf16 elsewhere, fp32 masters")
    ap.add_argument('--transport', default='nvlink', choices=['nvlink', 'nccl'])
    ap.add_argument('--variant', default='auto', choices=['auto', 'one_shot', 'two_shot', 'nvls'])
    ap.add_argument('--overlap', type=int, default=-1,
                    help='bucketed reduce+update launched during backward (-1: on when --gpus > 1)')
    ap.add_argument('--bucket-mb', type=float, default=4.0)
    ap.add_argument('--layout', default='auto', choices=['auto', 'contiguous', 'channels_last_3d'], help='R1 arm only')
    ap.add_argument('--native', type=int, default=1, help='use the hand-written sm_100a model kernels')
    ap.add_argument('--skip-e2e', action='store_true')
    ap.add_argument('--input-dtype', default='fp32', choices=['bf16', 'fp32'],
                    help='dtype of the volumes in pinned host memory for the headline e2e number (fp32 = what the '
                         'reference arm moves; the bf16-host variant is reported next to it as e2e_bf16_host)')
    ap.add_argument('--graph', type=int, default=1, help='capture the whole step in a CUDA graph')
    return ap.parse_args()


# ----------------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock + throttle reasons of one GPU, sampled every ~2 ms by an NVML thread while a timed region runs
    (a timed region of 20 steps is ~40 ms: an `nvidia-smi -lms` child cannot resolve that).  Falls back to one
    `nvidia-smi` query when NVML is not importable."""
    REASONS = {'hw_slowdown': 0x8, 'sw_power_cap': 0x4, 'sw_thermal_slowdown': 0x20, 'hw_thermal_slowdown': 0x40}

    def __init__(self, index):
        self.index, self.thread, self.stop_flag = index, None, False
        self.sm, self.mask, self.max_mhz, self.power = [], 0, None, []
        self.h = None
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            pr = torch.cuda.get_device_properties(index)
            bus = f'{getattr(pr, "pci_domain_id", 0):08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0'
            try:
                self.h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.nv = pynvml
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.h = None

    def _loop(self):
        nv = self.nv
        reasons = getattr(nv, 'nvmlDeviceGetCurrentClocksEventReasons', None) or \
            getattr(nv, 'nvmlDeviceGetCurrentClocksThrottleReasons', None)
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                if reasons is not None:
                    self.mask |= int(reasons(self.h))
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        self.sm, self.mask, self.power, self.stop_flag = [], 0, [], False
        if self.h is not None:
            import threading
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()

    def stop(self):
        out = {'sm_mhz': None, 'sm_max_mhz': self.max_mhz, 'reasons': [], 'samples': 0}
        if self.thread is not None:
            self.stop_flag = True
            self.thread.join(timeout=2)
            self.thread = None
        if self.sm:
            out.update(sm_mhz=statistics.median(self.sm), samples=len(self.sm),
                       power_w_max=max(self.power) if self.power else None)
            out['reasons'] = sorted(k for k, bit in self.REASONS.items() if self.mask & bit)
            return out
        try:    # fall-back: one query (after the region; better than nothing)
            q = 'clocks.sm,clocks.max.sm'
            r = subprocess.run(['nvidia-smi', f'--query-gpu={q}', '--format=csv,noheader,nounits', '-i', str(self.index)],
                               capture_output=True, text=True, timeout=10).stdout.split(',')
            out.update(sm_mhz=float(r[0]), sm_max_mhz=float(r[1]), samples=1)
        except Exception:
            pass
        return out


def merge_clocks(*parts):
    """One `clocks` record from the samplers of several timed regions (device-timed + end-to-end)."""
    sm = [p['sm_mhz'] for p in parts if p and p.get('sm_mhz')]
    out = {'sm_mhz': statistics.median(sm) if sm else None,
           'sm_max_mhz': max([p['sm_max_mhz'] for p in parts if p and p.get('sm_max_mhz')] or [None]),
           'reasons': sorted({r for p in parts if p for r in p.get('reasons', [])}),
           'samples': sum(p.get('samples', 0) for p in parts if p)}
    pw = [p.get('power_w_max') for p in parts if p and p.get('power_w_max')]
    if pw:
        out['power_w_max'] = max(pw)
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
    import importlib.util    # by path: the reference arm must not import our package
    spec = importlib.util.spec_from_file_location('_coinn_affinity', os.path.join(ROOT, 'coinstac_dinunet_b200', 'utils', 'affinity.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    aff = mod.pin_to_gpu(local, ranks_per_node=max(1, (n + 1) // 2))      # NUMA-local CPUs + pinned buffers, sane thread pool
    if not dist.is_initialized():
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    assert dist.get_world_size() == n, f'--gpus {n} but WORLD_SIZE={dist.get_world_size()} (launch with torchrun)'
    return dist.get_rank(), local, aff


def timed(fn, dist, torch):
    """barrier + sync | events around fn | sync + barrier; returns max-over-ranks milliseconds."""
    import gc
    gc.collect()
    gc.disable()                     # a generation-2 collection inside a 40 ms region is a 10 % outlier
    try:
        dist.barrier(device_ids=[torch.cuda.current_device()])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        result = fn()
        e1.record()
        torch.cuda.synchronize()
        dist.barrier(device_ids=[torch.cuda.current_device()])
    finally:
        gc.enable()
    ms = torch.tensor([e0.elapsed_time(e1)], device='cuda')
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item()), result


def r1_number(model, n):
    """samples/s of the R1 arm (cuDNN bf16 + NCCL + fused Adam, CUDA graph) recorded for this model at N GPUs."""
    try:
        with open(os.path.join(ROOT, 'baseline', 'r1_measured.json')) as fp:
            return float(json.load(fp)[model][str(n)]['value'])
    except Exception:
        return None


def metric_name(model):
    return 'samples/sec (whole box, max over sites) VBM-3D-CNN dSGD' if model == 'vbm' \
        else 'samples/sec (whole box, max over sites) FreeSurfer-MLP dSGD'


def gap_stats(stamps):
    """Host-side inter-step gaps (ms) of an end-to-end region: shows whether the host ever starved the device."""
    if not stamps or len(stamps) < 3:
        return None
    gaps = sorted((b - a) * 1e3 for a, b in zip(stamps[:-1], stamps[1:]))
    return {'p50': round(gaps[len(gaps) // 2], 3), 'max': round(gaps[-1], 3)}


# ------------------------------------------------------------------------------------- ours
def run_ours(a):
    import torch
    import torch.distributed as dist
    rank, local, aff = dist_setup(a.gpus)
    from coinstac_dinunet_b200 import ops
    from coinstac_dinunet_b200.engine import DistEngine
    from coinstac_dinunet_b200.models import FSVTrainer, InMemorySynthetic, VBMTrainer
    from coinstac_dinunet_b200.models.common import pinned_collate

    shape = VBM_SHAPE if a.model == 'vbm' else FS_SHAPE
    batch = a.batch or (8 if a.model == 'vbm' else 16)
    trainer_cls = VBMTrainer if a.model == 'vbm' else FSVTrainer
    n_distinct = 4 * batch                                   # 4 rotating batches: 272 MB fp32 for VBM (> L2)
    n_files = batch * (a.warmup + a.steps + 4) * 2
    overlap = (a.gpus > 1) if a.overlap < 0 else bool(a.overlap)
    host_dt = {'fp32': torch.float32, 'bf16': torch.bfloat16}

    class Volumes(InMemorySynthetic):
        def __init__(self, **kw):
            super().__init__(shape=shape, num_class=2, seed=100 + rank, pin=True,
                             dtype=host_dt[a.input_dtype] if a.model == 'vbm' else torch.float32, **kw)

    work = tempfile.mkdtemp(prefix='coinn_bench_') if rank == 0 else None
    box = [work]
    dist.broadcast_object_list(box, src=0)
    work = box[0]
    spec = dict(task_id='bench', mode='train', data_dir='data', num_class=2, batch_size=batch,
                split_ratio=[0.98, 0.01, 0.01], epochs=10 ** 6, gpus=[local], agg_engine='dSGD', seed=11,
                monitor_metric='f1', learning_rate=1e-3, validation_epochs=10 ** 9, transport=a.transport, reduce_variant=a.variant,
                compute_dtype='bf16' if a.dtype == 'fp8' else a.dtype, conv_backend='fp8' if a.dtype == 'fp8' else 'auto',
                channels_last='3d' if a.model == 'vbm' else None, native_ops=bool(a.native),
                input_shape=list(shape), input_size=shape[0], synthetic_distinct=n_distinct,
                overlap_backward=overlap, bucket_bytes=int(a.bucket_mb * (1 << 20)), prefetch_depth=3,
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

    def rounds(k, resident, readback, dtype=None):
        eng.cache['steps_per_round'] = k
        eng.cache['readback_per_step'] = readback
        eng.cache['synthetic_device'] = f'cuda:{local}' if resident else None
        # end-to-end mode: stage the next batches on a copy stream while the current step computes (public dataloader option)
        eng.cache['prefetch_to_device'] = None if resident else f'cuda:{local}'
        ds = eng.cache['dataset'].get('train')
        if ds is not None:
            want = host_dt[dtype] if (dtype and a.model == 'vbm') else ds.dtype
            if ds._x is None or ds._x.is_cuda != resident or ds.dtype != want:
                ds.dtype, ds._x = want, None                 # re-materialise on the requested side / in the requested dtype
                eng.cache['cursor'] = 0                       # ... and iterate the (re)placed data from the start
        # otherwise the site keeps streaming: the loader iterator (with the batches its prefetcher has staged) carries over
        # from the warm-up round into the timed round, so the timed region is steady state, not an epoch start
        # validation_epochs is huge in the spec, so the aggregator answers every finished round
        # ("epoch") with mode=train and the next round trains again
        eng.step(local_fn, remote_fn)

    sampler = ClockSampler(local)

    # ---- device-resident inputs: the kernel-side number ----
    rounds(a.warmup, True, False)
    l0 = ops.launch_count
    sampler.start()
    ms, _ = timed(lambda: rounds(a.steps, True, False), dist, torch)
    clocks = [sampler.stop()]
    launches = ops.launch_count - l0

    # ---- end to end: pinned host -> device every step, loss read back every step ----
    def e2e_run(dtype):
        rounds(3, False, True, dtype)                         # (re)materialise the pinned pool, allocate staging buffers
        rounds(max(a.warmup, 3), False, True, dtype)          # W untimed steps in exactly the timed configuration
        eng.cache['_host_stamps'] = stamps = []
        sampler.start()
        ms_e, _ = timed(lambda: rounds(a.steps, False, True, dtype), dist, torch)
        clocks.append(sampler.stop())
        eng.cache.pop('_host_stamps', None)
        item = 2 if (dtype == 'bf16' and a.model == 'vbm') else 4
        h2d = batch * (int(torch.tensor(shape).prod()) * item + 8)
        return {'value': a.gpus * batch * a.steps / (ms_e / 1e3), 'unit': 'samples/s', 'ms_per_step': ms_e / a.steps,
                'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': 4, 'host_input_dtype': dtype if a.model == 'vbm' else 'fp32',
                'losses_read_back': int(eng.cache.get('losses_read', 0)), 'host_step_gap_ms': gap_stats(stamps)}

    e2e = e2e_alt = None
    if not a.skip_e2e:
        e2e = e2e_run(a.input_dtype)
        if a.model == 'vbm':
            e2e_alt = e2e_run('bf16' if a.input_dtype == 'fp32' else 'fp32')

    arena = eng.cache['_arena']
    if rank == 0:
        value = a.gpus * batch * a.steps / (ms / 1e3)
        r1 = r1_number(a.model, a.gpus)
        ov = getattr(arena, '_overlap', None)
        line = {
            'metric': metric_name(a.model),
            'value': value, 'unit': 'samples/s', 'n_gpus': a.gpus, 'steps': a.steps, 'warmup': a.warmup,
            'ms_per_step': ms / a.steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': (value / r1) if r1 else None,
            'dtype': a.dtype, 'data': 'synthetic', 'impl': 'ours',
            'config': {'precision_recipe': ('MX-FP8 (e4m3 + ue8m0 per 32 channels, tcgen05 block_scale) for conv fprop/dgrad with C_in >= 32; '
                                            'bf16 first two blocks, wgrad and head; fp32 master weights in the fused reduce+Adam')
                       if a.dtype == 'fp8' else a.dtype,
                       'model': 'VBMNet 5x[Conv3d-BN-ReLU-MaxPool] + 3 FC, 3.55M params' if a.model == 'vbm'
                       else 'FSNet MLP 66-256-128-64-32-2',
                       'input': list(shape), 'per_site_batch': batch, 'global_batch': batch * a.gpus,
                       'parallelism': f'dSGD sites={a.gpus} (1 site/GPU)', 'optimizer': 'Adam(1e-3)',
                       'transport': arena.backend,
                       'reduce_variant': [arena._pick_variant(n * 4) for _, n in arena._launch_units()],
                       'overlap_backward': bool(ov), 'reduce_buckets': len(ov['buckets']) if ov else 1,
                       'native_model_kernels': bool(a.native), 'cuda_graph': bool(a.graph),
                       'host_input_dtype': a.input_dtype if a.model == 'vbm' else 'fp32',
                       'baseline_for_vs_baseline': 'R1 = same nn.Module, cuDNN/cuBLAS bf16 autocast + NCCL all_reduce(AVG) + '
                                                   'Adam(fused) in one CUDA graph, measured on B200 (baseline/r1_measured.json)',
                       'l2_policy': 'inputs rotate over 4 distinct batches; activations per step >> 126 MB L2',
                       'affinity': aff},
            'clocks': merge_clocks(*clocks), 'e2e': e2e, 'gpu_launches': launches,
        }
        if e2e_alt is not None:
            line['e2e_bf16_host' if a.input_dtype == 'fp32' else 'e2e_fp32_host'] = e2e_alt
        emit(line)
    dist.barrier(device_ids=[local])
    dist.destroy_process_group()


# --------------------------------------------------------------------------------------- R1
class _HostPool:
    """4 rotating batches in pinned host memory, handed out as zero-copy slices (what our arm's dataset does)."""

    def __init__(self, shape, batch, steps, seed, torch, resident_device=None):
        g = torch.Generator().manual_seed(seed)
        self.x = torch.randn((4 * batch, *shape), generator=g)
        self.y = torch.randint(0, 2, (4 * batch,), generator=g)
        if resident_device is not None:
            self.x, self.y = self.x.to(resident_device), self.y.to(resident_device)
        else:
            self.x, self.y = self.x.pin_memory(), self.y.pin_memory()
        self.batch, self.steps = batch, steps

    def __len__(self):
        return self.steps

    def __iter__(self):
        b = self.batch
        for i in range(self.steps):
            j = (i % 4) * b
            yield {'inputs': self.x[j:j + b], 'labels': self.y[j:j + b]}


def run_r1(a):
    import torch
    import torch.distributed as dist
    rank, local, aff = dist_setup(a.gpus)
    sys.path.insert(0, os.path.join(ROOT, 'baseline'))
    from r1_nccl_cudnn import R1Step, launch_list
    from coinstac_dinunet_b200.data.data import DevicePrefetcher
    from coinstac_dinunet_b200.parallel.nvlink_learner import _LaggedReadback
    torch.backends.cudnn.benchmark = True
    dev = torch.device('cuda', local)
    shape = VBM_SHAPE if a.model == 'vbm' else FS_SHAPE
    batch = a.batch or (8 if a.model == 'vbm' else 16)
    step = R1Step(a.model, shape, batch, dev, lr=1e-3, layout=a.layout, graph=bool(a.graph))
    sampler = ClockSampler(local)

    def run(k, resident, readback):
        pool = _HostPool(shape, batch, k, 100 + rank, torch, resident_device=dev if resident else None)
        it = pool if resident else DevicePrefetcher(pool, dev, depth=3)
        rb = _LaggedReadback(dev) if readback else None

        def go():
            for b in it:
                loss = step.step(b['inputs'], b['labels'])
                if rb is not None:
                    rb.push(loss)
            if rb is not None:
                rb.finish()
        return go

    run(a.warmup, True, False)()
    go = run(a.steps, True, False)
    sampler.start()
    ms, _ = timed(go, dist, torch)
    clocks = [sampler.stop()]
    e2e = None
    if not a.skip_e2e:
        run(max(a.warmup, 3), False, True)()
        go = run(a.steps, False, True)
        sampler.start()
        ms_e, _ = timed(go, dist, torch)
        clocks.append(sampler.stop())
        h2d = batch * (int(torch.tensor(shape).prod()) * 4 + 8)
        e2e = {'value': a.gpus * batch * a.steps / (ms_e / 1e3), 'unit': 'samples/s', 'ms_per_step': ms_e / a.steps,
               'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': 4, 'host_input_dtype': 'fp32'}
    top = None
    if rank == 0 and os.environ.get('COINN_R1_LAUNCHES'):
        top = [[k, c, round(t, 1)] for k, c, t in launch_list(step)[:25]]
    if rank == 0:
        value = a.gpus * batch * a.steps / (ms / 1e3)
        line = {'metric': metric_name(a.model), 'value': value, 'unit': 'samples/s', 'n_gpus': a.gpus, 'steps': a.steps,
                'warmup': a.warmup, 'ms_per_step': ms / a.steps, 'higher_is_better': True, 'scaling': 'weak',
                'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic', 'impl': 'nccl_cudnn',
                'config': {'model': 'RefVBMNet' if a.model == 'vbm' else 'RefFSNet', 'input': list(shape),
                           'per_site_batch': batch, 'global_batch': batch * a.gpus,
                           'parallelism': f'data parallel sites={a.gpus}: flat-grad NCCL all_reduce(AVG) + Adam(fused, capturable)',
                           'memory_format': step.layout, 'autocast': 'bf16', 'cudnn_benchmark': True,
                           'cuda_graph': bool(a.graph), 'affinity': aff},
                'clocks': merge_clocks(*clocks), 'e2e': e2e, 'gpu_launches': 0}
        if top:
            line['top_kernels_us'] = top
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
    rank, local, _aff = dist_setup(a.gpus)
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
            'metric': metric_name(a.model),
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
    elif args.impl == 'nccl_cudnn':
        run_r1(args)
    else:
        run_ours(args)
