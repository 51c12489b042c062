"""Cross-GPU correctness of the fused NVLink kernels and of the full protocol on >= 2 GPUs."""
import pytest
import torch

import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from test_dist_cpu import run_workers  # noqa: E402

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]


def _n():
    return min(torch.cuda.device_count(), 8)


def test_fused_reduce_all_variants(tmp_path):
    res = run_workers('fused', tmp_path, nproc=_n(), port=29701)
    for r in res['results']:
        assert r['identical'], r
        assert r['zeroed'], r
        assert r['err'] < 5e-6, r


def test_protocol_over_nvlink(tmp_path):
    res = run_workers('protocol', tmp_path, nproc=_n(), port=29702, extra=['transport=nvlink'])
    assert res['backend'] == 'nvlink' and res['replicas_identical'] and res['csv']


def test_symm_allreduce_matches_nccl(tmp_path):
    res = run_workers('allreduce', tmp_path, nproc=_n(), port=29703)
    assert all(r['err'] < 1e-5 for r in res['results']), res


def test_powersgd_protocol_over_nvlink(tmp_path):
    res = run_workers('protocol', tmp_path, nproc=_n(), port=29704, extra=['transport=nvlink', 'agg_engine=powerSGD'])
    assert res['backend'] == 'nvlink' and res['csv'] and res['trace'][-2] == 'success'


def test_bucketed_overlap_matches_single_launch(tmp_path):
    res = run_workers('overlap', tmp_path, nproc=_n(), port=29705)
    for r in res['results']:
        assert r['how'] == 'bucketed' and r['buckets'] >= 3 and r['steps'] == 6, r
        assert r['err'] < 1e-6 and r['identical'] and r['zeroed'], r
