"""Multi-process (gloo, world_size 2) tests of the one-process-per-site engine and of the host
logic of the NVLink learners (the data plane falls back to all-reduce + optimizer.step on CPU)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_workers(scenario, work, nproc=2, port=29611, extra=(), timeout=600, env=None):
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={nproc}',
           '--master-addr', '127.0.0.1', '--master-port', str(port),
           os.path.join(ROOT, 'tests', 'dist_worker.py'), scenario, str(work), *extra]
    e = dict(os.environ)
    e.update(env or {})
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=e)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    with open(os.path.join(work, 'result.json')) as fp:
        return json.load(fp)


@pytest.mark.parametrize('engine,port', [('dSGD', 29611), ('powerSGD', 29612), ('rankDAD', 29613)])
def test_dist_engine_two_sites(tmp_path, engine, port):
    res = run_workers('protocol', tmp_path, port=port, extra=[f'agg_engine={engine}'])
    assert res['csv'] and res['trace'][-2] == 'success'
    assert res['replicas_identical']
    assert res['backend'] == 'torch' and res['fused_steps'] > 0


def test_pretrained_weights_broadcast_without_files_on_device_transports(tmp_path):
    """C5: on the nvlink / nccl transports only the pre-training site reads its checkpoint; the other sites receive
    parameters, buffers and optimizer state through DistArena.broadcast_from (gloo stand-in here, NVLink peer copies on
    GPUs) and all replicas stay identical through the rest of the run."""
    res = run_workers('protocol', tmp_path, port=29614, extra=['agg_engine=dSGD', 'pretrain=1'])
    assert res['csv'] and res['trace'][-2] == 'success' and res['replicas_identical']
    assert 'pre_computation' in [str(t).split('.')[-1].lower() for t in res['trace']]
    assert res['weights_broadcast'] == 'device'


def test_control_plane_mailbox_handles_oversized_messages(tmp_path):
    """Per-round JSON goes through the shared-memory mailbox; a message that does not fit a slot (here: 256-byte slots, so
    nearly every message) transparently travels point-to-point over torch.distributed instead - same run, same result."""
    res = run_workers('protocol', tmp_path, port=29615, extra=['agg_engine=dSGD'], env={'COINN_CTL_SLOT_BYTES': '256'})
    assert res['csv'] and res['trace'][-2] == 'success' and res['replicas_identical']
    res2 = run_workers('protocol', tmp_path / 'gloo', port=29616, extra=['agg_engine=dSGD'], env={'COINN_CTL_SHM': '0'})
    assert res2['trace'] == res['trace'] and res2['rounds'] == res['rounds']      # (the fold seed is drawn per run)


def test_epoch_level_resume_over_the_device_transport(tmp_path):
    """The resume points of ``checkpoint_epochs`` through DistEngine + NvlinkLearner (one engine round = one epoch of fused
    steps): a 2-site run killed after the epoch-3 point was committed and restarted with ``resume=1`` ends with the weights
    and the training log of the undisturbed run."""
    common = ['agg_engine=dSGD', 'epochs=5', 'checkpoint_epochs=1']
    clean = run_workers('protocol', tmp_path / 'clean', port=29617, extra=common)
    cut = run_workers('protocol', tmp_path / 'cut', port=29618, extra=common + ['die_at_epoch=3'])
    assert cut == {'died_at_epoch': 3}
    res = run_workers('protocol', tmp_path / 'cut', port=29619, extra=common + ['resume=1'])
    assert res['resumed_epoch'] == 3 and res['trace'][-2] == 'success' and res['replicas_identical']
    assert res['param_sum'] == clean['param_sum'] and res['train_log'] == clean['train_log']
    assert clean['resumed_epoch'] is None
