"""Worker launched by torch.distributed.run from the tests (gloo on CPU, nccl on GPUs).

usage: dist_worker.py <scenario> <work_dir> [key=value ...]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def scenario_protocol(work, opts):
    """Full k-fold dSGD run with one process per site through DistEngine + NvlinkLearner."""
    from coinstac_dinunet_b200.engine import DistEngine
    from coinstac_dinunet_b200.models import FSVDataset, FSVTrainer, write_synthetic_site
    spec = dict(task_id='fsv', mode='train', data_dir='data', labels_file='labels.json', input_size=66, num_class=2,
                batch_size=4, num_folds=None, split_ratio=[0.6, 0.2, 0.2], learning_rate=1e-2, seed=7,
                transport=opts.get('transport', 'nvlink'), agg_engine=opts.get('agg_engine', 'dSGD'),
                start_powerSGD_iter=2, matrix_approximation_rank=2, cuda_graph=opts.get('cuda_graph') == '1',
                epochs=int(opts.get('epochs', 2)), reduce_variant=opts.get('reduce_variant', 'auto'),
                overlap_backward=opts.get('overlap') == '1', bucket_bytes=int(opts.get('bucket_bytes', 64 << 10)),
                precision_bits=int(opts.get('precision_bits', 32)), dad_reduction_rank=int(opts.get('dad_rank', 10)),
                checkpoint_epochs=int(opts.get('checkpoint_epochs', 0)), resume=opts.get('resume') == '1',
                gpus=[int(os.environ.get('LOCAL_RANK', 0))] if torch.cuda.is_available() else None)
    eng = DistEngine(work, inputspec=spec)
    died = {}
    if opts.get('die_at_epoch'):                        # power cut once the aggregator has committed that resume point
        class PowerCut(Exception):
            pass
        orig_step = eng.step

        def step(lf, rf):
            ok = orig_step(lf, rf)
            box = [None]
            if eng.rank == 0:
                path = os.path.join(eng.remote_state['outputDirectory'], 'fsv', 'resume.json')
                if os.path.exists(path):
                    with open(path) as fp:
                        box[0] = (json.load(fp).get('in_progress') or {}).get('epoch')
            dist.broadcast_object_list(box, src=0)
            if box[0] is not None and box[0] >= int(opts['die_at_epoch']):
                died['epoch'] = box[0]
                raise PowerCut()
            return ok
        eng.step = step
    # count torch.distributed collectives issued INSIDE a compressed / rankDAD optimizer step (the device data plane
    # must not issue any: the exchange is in-kernel over symmetric memory)
    import coinstac_dinunet_b200.parallel.nvlink_learner as nvl
    inside = {'flag': False, 'calls': 0, 'steps': 0}
    for fn_name in ('all_reduce', 'all_gather', 'broadcast', 'all_gather_into_tensor', 'reduce_scatter_tensor', 'all_to_all_single'):
        orig = getattr(dist, fn_name, None)
        if orig is None:
            continue

        def counted(*a, _orig=orig, **kw):
            if inside['flag']:
                inside['calls'] += 1
            return _orig(*a, **kw)
        setattr(dist, fn_name, counted)
        setattr(nvl._dist, fn_name, counted)
    for cls, meth in ((nvl.NvlinkPowerSGDLearner, '_compressed_step'), (nvl.NvlinkDADLearner, '_dad_step')):
        orig_m = getattr(cls, meth)

        def wrapped(self, _orig=orig_m):
            inside['flag'] = True
            inside['steps'] += 1
            try:
                return _orig(self)
            finally:
                inside['flag'] = False
        setattr(cls, meth, wrapped)
    sizes = [24, 18, 30, 12, 20, 16, 28, 22]
    write_synthetic_site(eng.state['baseDirectory'], sizes[eng.rank % len(sizes)], (66,), seed=eng.rank)
    local_kw = {'pretrain_args': {'epochs': 2}} if opts.get('pretrain') == '1' else None
    try:
        rounds = eng.run_nodes(FSVTrainer, FSVDataset, local_kw=local_kw, remote_kw={'seed': 7}, max_rounds=500)
    except Exception:
        if not died:
            raise
        if eng.rank == 0:
            with open(os.path.join(work, 'result.json'), 'w') as fp:
                json.dump({'died_at_epoch': died['epoch']}, fp)
        return
    model = eng.cache['nn']['fs_net']
    flat = torch.cat([p.detach().float().reshape(-1).cpu() for p in model.parameters()])
    gathered = [None] * eng.world
    dist.all_gather_object(gathered, flat)
    if eng.rank == 0:
        same = all(torch.equal(gathered[0], g) for g in gathered[1:])
        csv = os.path.join(eng.remote_state['outputDirectory'], 'fsv', 'global_test_metrics.csv')
        res = {'rounds': rounds, 'replicas_identical': bool(same), 'csv': os.path.exists(csv),
               'backend': eng.cache['_arena'].backend, 'fused_steps': eng.cache['_arena'].steps_done,
               'graphed': '_graph_step' in eng.cache, 'param_sum': float(gathered[0].double().sum()),
               'trace': [t['remote'] for t in eng.trace], 'weights_broadcast': eng.cache.get('_weights_broadcast'),
               'collectives_in_steps': inside['calls'], 'compressed_steps': inside['steps'],
               'resumed_epoch': eng.remote_cache.get('resumed_epoch'), 'train_log': eng.remote_cache.get('train_log')}
        with open(os.path.join(work, 'result.json'), 'w') as fp:
            json.dump(res, fp)


SCENARIOS = {'protocol': scenario_protocol}

if __name__ == '__main__':
    name, work = sys.argv[1], sys.argv[2]
    opts = dict(kv.split('=', 1) for kv in sys.argv[3:])
    from coinstac_dinunet_b200.engine import init_process_group
    init_process_group()
    if torch.cuda.is_available():
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        import dist_scenarios_gpu  # registers the GPU scenarios
        SCENARIOS.update(dist_scenarios_gpu.SCENARIOS)
    SCENARIOS[name](work, opts)
    dist.barrier()
    dist.destroy_process_group()
