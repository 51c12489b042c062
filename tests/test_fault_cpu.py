"""Fault injection: a site that dies mid-epoch surfaces as an exception carrying the
node's last output; retrying the round does not double-apply the sites that had already finished it."""
import pytest
import torch

from coinstac_dinunet_b200.models import FSVDataset, FSVTrainer


def _flat(cache):
    return torch.cat([p.detach().reshape(-1) for p in cache['nn']['fs_net'].parameters()])


def test_site_failure_and_round_retry(fs_sites):
    spec = {'num_folds': None, 'split_ratio': [0.6, 0.2, 0.2], 'epochs': 2}
    clean = fs_sites(spec=spec)
    clean.run_nodes(FSVTrainer, FSVDataset, max_rounds=1000)
    want = _flat(clean.site_cache['local0'])

    eng = fs_sites(spec=spec)
    fired = []

    def hook(rnd, site):
        if rnd == 7 and site == 'local1' and not fired:      # local0 has already computed round 7
            fired.append(rnd)
            raise RuntimeError('injected: site local1 lost')
    eng.fault_hook = hook
    with pytest.raises(RuntimeError, match='injected'):
        eng.run_nodes(FSVTrainer, FSVDataset, max_rounds=1000)
    assert eng.round == 7 and 'local0' in eng._partial
    eng.run_nodes(FSVTrainer, FSVDataset, max_rounds=1000)     # retry: continues from the failed round
    assert eng.trace[-2]['remote'] == 'success'
    assert torch.equal(_flat(eng.site_cache['local0']), want), 'retry must be equivalent to an undisturbed run'
    assert torch.equal(_flat(eng.site_cache['local0']), _flat(eng.site_cache['local1']))


def test_node_exception_carries_last_output(fs_sites):
    """COINNLocal.__call__ re-raises with ``self.out`` attached (ref local.py:289-295)."""
    eng = fs_sites(spec={'num_folds': None, 'split_ratio': [0.6, 0.2, 0.2], 'epochs': 1})

    class Broken(FSVTrainer):
        def iteration(self, batch):
            raise ValueError('boom')
    with pytest.raises(Exception) as exc:
        eng.run_nodes(Broken, FSVDataset, max_rounds=50)
    assert 'phase' in str(exc.value)


def test_fold_level_resume(fs_sites, tmp_path):
    """A run that dies after fold 0 is restarted with fresh node caches and ``resume=True``: finished folds are
    skipped, the final aggregate still covers all three folds."""
    import os
    spec = {'num_folds': 3, 'epochs': 1}
    eng = fs_sites(spec=spec)

    def stop_after_first_fold(rnd, site):
        if sum(t['remote'] == 'next_run' for t in eng.trace) >= 2:      # fold 0 done, fold 1 announced
            raise KeyboardInterrupt('power cut')
    eng.fault_hook = stop_after_first_fold
    with pytest.raises(KeyboardInterrupt):
        eng.run_nodes(FSVTrainer, FSVDataset, max_rounds=2000)
    resume_file = os.path.join(eng.remote_state['outputDirectory'], 'fsv', 'resume.json')
    assert os.path.exists(resume_file)

    from coinstac_dinunet_b200.engine import InProcessEngine
    base = dict(task_id='fsv', mode='train', data_dir='data', labels_file='labels.json', input_size=66, num_class=2,
                batch_size=4, learning_rate=1e-2, seed=7, monitor_metric='f1', metric_direction='maximize',
                log_header='Loss|Accuracy,F1', verbose=False, **spec)
    eng2 = InProcessEngine(eng.work_dir, n_sites=2, inputspec=base)      # same directories, empty caches
    eng2.run_nodes(FSVTrainer, FSVDataset, remote_kw={'resume': True}, max_rounds=2000)
    assert eng2.remote_cache['resumed_folds'] == ['0']
    assert [t['remote'] for t in eng2.trace].count('next_run') == 2      # only folds 1 and 2 were trained
    rows = open(os.path.join(eng2.remote_state['outputDirectory'], 'fsv', 'global_test_metrics.csv')).read().strip().split('\n')
    assert len(rows) == 2 and len(eng2.remote_cache['serializable_global_test_scores']) == 3


_MAILBOX_PEER = """
import os, sys
from coinstac_dinunet_b200.engine.shm_plane import ShmMailbox
mb = ShmMailbox(sys.argv[1], 1, 2, slot_bytes=4096, create=False)      # rank 1 of 2
mb.gather({'round': 1})
assert mb.broadcast() == 'ack'
if sys.argv[2] == 'abort':                              # what DistEngine.step does when a node raises
    mb.abort()
    os._exit(0)
os._exit(17)                                            # no abort word, no clean-up - like a segfault or an OOM kill
"""


@pytest.mark.parametrize('behaviour', ['abort', 'die', 'zombie'])
def test_control_plane_mailbox_detects_a_failed_peer(behaviour):
    """A rank waiting on the shared-memory control plane learns within a fraction of a second that the rank it waits for
    raised (abort word) or vanished (pid check; a dead-but-unreaped process counts as gone) - not after the half-hour
    timeout."""
    import os
    import subprocess
    import sys
    import time
    from coinstac_dinunet_b200.engine.shm_plane import PeerFailure, ShmMailbox
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    mb = ShmMailbox(None, 0, 2, slot_bytes=4096, create=True, timeout_s=60.0)
    proc = None
    try:
        proc = subprocess.Popen([sys.executable, '-c', _MAILBOX_PEER, mb.name, behaviour], cwd=root)
        assert mb.gather({'round': 1}) == [{'round': 1}, {'round': 1}]
        mb.broadcast('ack')
        if behaviour == 'zombie':                       # dead but not reaped by its launcher yet: /proc/<pid>/stat says Z
            deadline = time.monotonic() + 60
            while open(f'/proc/{proc.pid}/stat').read().rsplit(')', 1)[1].split()[0] != 'Z':
                assert time.monotonic() < deadline, 'the peer process never exited'
                time.sleep(0.01)
        else:
            proc.wait(30)                               # reaped: the pid is really gone
        t0 = time.monotonic()
        with pytest.raises(PeerFailure) as err:
            mb.gather({'round': 2})                     # rank 1 never posts message 2
        assert time.monotonic() - t0 < 5.0
        assert ('aborted' if behaviour == 'abort' else 'is gone') in str(err.value)
    finally:
        if proc is not None:
            proc.wait(30)
        mb.close()


@pytest.mark.parametrize('engine', ['dSGD', 'powerSGD', 'rankDAD'])
def test_epoch_level_resume_continues_a_fold_bit_exactly(fs_sites, engine):
    """``checkpoint_epochs=1``: after every validation round the aggregator requests a resume point, the sites write
    ``resume.<task>-<fold>.e<epoch>.pt`` and the point is committed once all of them answered.  A run killed in the middle of
    a fold and restarted with empty caches and ``resume=True`` picks the fold up at the committed epoch and finishes with
    exactly the weights, logs and scores of an undisturbed run.  With PowerSGD the point also carries the engine state
    (error feedback, warm-start factors, iteration count)."""
    import json
    import os
    from coinstac_dinunet_b200.engine import InProcessEngine
    spec = {'num_folds': None, 'split_ratio': [0.6, 0.2, 0.2], 'epochs': 5, 'checkpoint_epochs': 1, 'agg_engine': engine,
            'start_powerSGD_iter': 2, 'matrix_approximation_rank': 2}
    clean = fs_sites(spec=spec)
    clean.run_nodes(FSVTrainer, FSVDataset, max_rounds=5000)
    want = _flat(clean.site_cache['local0'])
    want_log = clean.remote_cache['train_log']

    import shutil
    shutil.rmtree(clean.work_dir)
    eng = fs_sites(spec=spec)
    seen = []

    def power_cut(rnd, site):
        path = os.path.join(eng.remote_state['outputDirectory'], 'fsv', 'resume.json')
        if os.path.exists(path):
            with open(path) as fp:
                point = json.load(fp).get('in_progress')
            if point and point['epoch'] >= 3:
                seen.append(point['epoch'])
                raise KeyboardInterrupt('power cut')
    eng.fault_hook = power_cut
    with pytest.raises(KeyboardInterrupt):
        eng.run_nodes(FSVTrainer, FSVDataset, max_rounds=5000)
    assert seen == [3]
    log_dir = os.path.join(eng.site_state['local0']['outputDirectory'], 'fsv', 'fold_0')
    kept = sorted(n for n in os.listdir(log_dir) if n.startswith('resume.'))
    assert 'resume.fsv-0.e3.pt' in kept and 'resume.fsv-0.e1.pt' not in kept       # older points are pruned

    base = dict(task_id='fsv', mode='train', data_dir='data', labels_file='labels.json', input_size=66, num_class=2,
                batch_size=4, learning_rate=1e-2, seed=7, monitor_metric='f1', metric_direction='maximize',
                log_header='Loss|Accuracy,F1', verbose=False, resume=True, **spec)
    eng2 = InProcessEngine(eng.work_dir, n_sites=2, inputspec=base)                 # same directories, empty caches
    eng2.run_nodes(FSVTrainer, FSVDataset, max_rounds=5000)
    assert eng2.remote_cache['resumed_epoch'] == 3
    assert eng2.trace[-2]['remote'] == 'success'
    assert torch.equal(_flat(eng2.site_cache['local0']), want), 'a resumed fold must end where the undisturbed run ends'
    assert torch.equal(_flat(eng2.site_cache['local0']), _flat(eng2.site_cache['local1']))
    assert eng2.remote_cache['train_log'] == want_log
