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
