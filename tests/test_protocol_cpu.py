"""Milestone A (BASELINE.json config 1): FreeSurfer MLP dSGD, 2 CPU sites, file transport,
all phases, k folds, artefacts."""
import glob
import os

import numpy as np
import pytest
import torch

from coinstac_dinunet_b200.models import FSVDataset, FSVTrainer


def _params(cache):
    model = cache['nn']['fs_net']
    return torch.cat([p.detach().reshape(-1) for p in model.parameters()])


def test_dsgd_two_cpu_sites_full_run(fs_sites):
    eng = fs_sites()
    rounds = eng.run_nodes(FSVTrainer, FSVDataset, max_rounds=2000)
    assert rounds > 10
    phases = [t['remote'] for t in eng.trace]
    assert phases[0] == 'next_run' and phases[-2] == 'success'  # last round only delivers the zip
    assert phases.count('next_run') == 3  # three folds
    # replicas stay bit-identical in parameters (buffers are deliberately local, quirk 15)
    a, b = (_params(eng.site_cache[s]) for s in eng.site_ids)
    assert torch.equal(a, b)

    out0 = eng.site_state['local0']['outputDirectory']
    for f in range(3):
        fold = os.path.join(out0, 'fsv', f'fold_{f}')
        assert os.path.exists(os.path.join(fold, f'latest.fsv-{f}.pt'))
        assert os.path.exists(os.path.join(fold, 'logs.json'))
    assert len(glob.glob(os.path.join(out0, 'fsv', 'splits', 'SPLIT_*.json'))) == 3
    rout = eng.remote_state['outputDirectory']
    assert os.path.exists(os.path.join(rout, 'fsv', 'global_test_metrics.csv'))
    assert os.path.exists(os.path.join(rout, 'fsv', 'fold_0', 'test_metrics.csv'))
    assert glob.glob(os.path.join(out0, 'fsv_dSGD_*.zip'))


def test_round_trace_matches_reference_table(fs_sites):
    """init_runs -> next_run -> computation... ; lagging site keeps training."""
    eng = fs_sites(spec={'num_folds': None, 'split_ratio': [0.6, 0.2, 0.2], 'epochs': 1})
    eng.run_nodes(FSVTrainer, FSVDataset, max_rounds=500)
    t = eng.trace
    assert t[0]['sites']['local0'][0] == 'init_runs' and t[0]['remote'] == 'next_run'
    assert t[1]['sites']['local0'][0] == 'computation'          # first to_reduce happens in the next_run call
    # local1 has fewer samples: it reaches validation_waiting first and keeps going
    first_wait = next(i for i, r in enumerate(t) if r['sites']['local1'][1] == 'validation_waiting')
    assert t[first_wait]['sites']['local0'][1] == 'train'
    assert t[first_wait + 1]['sites']['local1'][0] == 'computation'
    assert any(all(m == 'validation' for m in r['modes'].values()) for r in t if r['modes'])
    assert any(all(m == 'test' for m in r['modes'].values()) for r in t if r['modes'])
    assert t[-2]['remote'] == 'success'


@pytest.mark.parametrize('engine', ['powerSGD', 'rankDAD'])
def test_compressed_engines_run(fs_sites, engine):
    spec = {'agg_engine': engine, 'num_folds': None, 'split_ratio': [0.6, 0.2, 0.2], 'epochs': 2,
            'start_powerSGD_iter': 2, 'matrix_approximation_rank': 2}
    eng = fs_sites(spec=spec)
    eng.run_nodes(FSVTrainer, FSVDataset, max_rounds=2000)
    assert eng.trace[-2]['remote'] == 'success'
    a, b = (_params(eng.site_cache[s]) for s in eng.site_ids)
    assert torch.allclose(a, b, atol=1e-5)


def test_training_learns(fs_sites):
    eng = fs_sites(sizes=(64, 48), spec={'num_folds': None, 'split_ratio': [0.7, 0.15, 0.15], 'epochs': 6,
                                         'learning_rate': 5e-3})
    eng.run_nodes(FSVTrainer, FSVDataset, max_rounds=5000)
    rows = open(os.path.join(eng.remote_state['outputDirectory'], 'fsv', 'global_test_metrics.csv')).read().split('\n')
    loss, acc, f1 = (float(v) for v in rows[1].split(',')[:3])
    assert acc > 0.7, rows


def test_half_precision_wire_format(fs_sites):
    """``precision_bits=16`` (reference local.py:54, learner.py:17): gradients travel as float16 object arrays; the
    replicas still end bit-identical because both sites apply the same averaged float16 gradient."""
    from coinstac_dinunet_b200.utils import tensorutils
    seen = []
    orig = tensorutils.save_arrays

    def spy(path, arrays):
        seen.append({str(np.asarray(a).dtype) for a in arrays})
        return orig(path, arrays)
    tensorutils.save_arrays = spy
    try:
        eng = fs_sites(spec={'num_folds': None, 'split_ratio': [0.6, 0.2, 0.2], 'epochs': 1, 'precision_bits': 16})
        eng.run_nodes(FSVTrainer, FSVDataset, max_rounds=500)
    finally:
        tensorutils.save_arrays = orig
    assert eng.trace[-2]['remote'] == 'success'
    assert seen and all(d == {'float16'} for d in seen), seen[:3]
    a, b = (_params(eng.site_cache[s]) for s in eng.site_ids)
    assert torch.equal(a, b)
