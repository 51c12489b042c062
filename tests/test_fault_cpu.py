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
