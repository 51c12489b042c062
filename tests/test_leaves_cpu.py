"""Unit tests of the leaf libraries: golden values from SURVEY §8.4, metric
oracles from NumPy / scikit-learn, checkpoint layout (§5.4), arg-merge priority (local.py:93-109)."""
import json
import os

import numpy as np
import pytest
import torch

from coinstac_dinunet_b200 import COINNLocal, config
from coinstac_dinunet_b200.config.keys import AGG_Engine, Key, Mode, Phase
from coinstac_dinunet_b200.data import COINNPaddedDataSampler, datautils
from coinstac_dinunet_b200.metrics import AUCROCMetrics, COINNAverages, ConfusionMatrix, Prf1a, dice_loss_binary
from coinstac_dinunet_b200.utils import FrozenDict, lazy_debug, performance_improved_, save_cache, save_scores, stop_training_
from coinstac_dinunet_b200.utils import tensorutils as tu


def test_enums_are_wire_strings():
    assert Phase.INIT_RUNS == 'init_runs' and json.dumps({'p': Phase.COMPUTATION}) == '{"p": "computation"}'
    assert Mode.VALIDATION_WAITING == 'validation_waiting' and f'{AGG_Engine.dSGD}' == 'dSGD'
    assert Key.TRAIN_SERIALIZABLE == 'serializable_train_scores'
    assert config.grads_file == 'grads.npy' and config.avg_grads_file == 'avg_grads.npy'
    assert config.boolean_string(' True ') and not config.boolean_string('no')


def test_lazy_debug_cadence():
    assert [x for x in range(1, 61) if lazy_debug(x)] == \
        [1, 2, 4, 6, 9, 12, 15, 18, 20, 24, 28, 32, 36, 40, 44, 48, 52, 55, 60]


def test_frozen_dict():
    d = FrozenDict({'a': 1})
    d['b'] = 2
    d.update(c=3)
    with pytest.raises(ValueError):
        d['a'] = 5
    with pytest.raises(ValueError):
        d.update(b=0)
    assert dict(d) == {'a': 1, 'b': 2, 'c': 3}


def test_ratio_and_kfold_splits_golden(tmp_path):
    files = [f'f{i}' for i in range(10)]
    sp = datautils.create_ratio_split(list(files), {'split_ratio': [.6, .2, .2], 'split_dir': None}, shuffle_files=False)
    assert sp == {'train': files[:6], 'validation': files[6:8], 'test': files[8:]}
    sp = datautils.create_ratio_split(list(files), {'split_ratio': [.8, .2], 'split_dir': None}, shuffle_files=False)
    assert sp == {'train': files[:8], 'test': files[8:]} and 'validation' not in sp
    cache = {'num_folds': 5, 'split_dir': str(tmp_path)}
    datautils.create_k_fold_splits(list(files), cache, shuffle_files=False)
    s0 = json.load(open(tmp_path / 'SPLIT_0.json'))
    s4 = json.load(open(tmp_path / 'SPLIT_4.json'))
    assert s0 == {'train': files[4:], 'validation': ['f2', 'f3'], 'test': ['f0', 'f1']}
    assert s4 == {'train': files[2:8], 'validation': ['f0', 'f1'], 'test': ['f8', 'f9']}
    shuffled = list(files)
    datautils._seeded_shuffle(shuffled)
    assert shuffled == ['f5', 'f2', 'f7', 'f1', 'f8', 'f4', 'f3', 'f6', 'f0', 'f9']


def test_init_k_folds_priority(tmp_path):
    state = {'baseDirectory': str(tmp_path / 'in'), 'outputDirectory': str(tmp_path / 'out')}
    os.makedirs(state['baseDirectory'])
    cache = {'task_id': 't', 'num_folds': 3}
    datautils.init_k_folds([f'f{i}' for i in range(9)], cache, state)
    assert cache['splits'] == {'0': 'SPLIT_0.json', '1': 'SPLIT_1.json', '2': 'SPLIT_2.json'}
    cache2 = {'task_id': 't2'}
    datautils.init_k_folds([], cache2, state)
    assert list(cache2['splits'].values()) == ['empty_split.json']


def test_padded_sampler_golden():
    assert list(COINNPaddedDataSampler(range(10), 4)) == [*range(10), 0, 1]
    assert list(COINNPaddedDataSampler(range(3), 8)) == [0, 1, 2, 0, 1, 2, 0, 1]
    s = COINNPaddedDataSampler(range(10), 4, seed=3, shuffle=True)
    a = list(s)
    s.set_epoch(1)
    assert sorted(a[:10]) == list(range(10)) and len(a) == 12 and list(s) != a
    assert len(COINNPaddedDataSampler(range(10), 4, drop_last=True)) == 8


def test_model_selection_helpers():
    cache = {'metric_direction': 'maximize', 'best_val_score': 0.5, 'best_val_epoch': 0, 'epochs': 10, 'patience': 2}
    assert not performance_improved_(1, 0.50005, cache)
    assert performance_improved_(2, 0.6, cache) and cache['best_val_epoch'] == 2
    assert not stop_training_(4, cache) and stop_training_(5, cache)
    cache = {'metric_direction': 'minimize', 'best_val_score': 1.0, 'best_val_epoch': 0, 'epochs': 3}
    assert performance_improved_(1, 0.9, cache) and not performance_improved_(2, 0.95, cache)


def test_averages():
    a = COINNAverages(num_averages=2)
    a.add(2.0, 4, 0); a.add(torch.tensor(1.0), 2, 0); a.add(3.0, 1, 1)
    assert a.get().tolist() == [round(10 / 6, 5), 3.0]
    b = COINNAverages(num_averages=2)
    b.reduce_sites([a.serialize(), a.serialize()])
    assert b.get().tolist() == a.get().tolist() and b.counts.tolist() == [12.0, 2.0]
    assert COINNAverages().get().tolist() == [0.0]


def test_prf1a_against_sklearn():
    from sklearn import metrics as skm
    g = torch.Generator().manual_seed(0)
    pred, true = torch.randint(0, 2, (500,), generator=g), torch.randint(0, 2, (500,), generator=g)
    m = Prf1a()
    m.add(pred[:200], true[:200]); m.add(pred[200:] * 255, true[200:] * 255)   # 255 == 1
    assert abs(m.precision - skm.precision_score(true, pred)) < 1e-4
    assert abs(m.recall - skm.recall_score(true, pred)) < 1e-4
    assert abs(m.f1 - skm.f1_score(true, pred)) < 1e-4
    assert abs(m.accuracy - skm.accuracy_score(true, pred)) < 1e-4
    assert abs(m.overlap - skm.jaccard_score(true, pred)) < 1e-4
    assert m.tp + m.fp + m.tn + m.fn == 500
    r = Prf1a()
    r.reduce_sites([m.serialize(), [1.0, 1.0, 1.0]])
    assert r.accuracy == round((m.accuracy + 1) / 2, 5) and r.extract('f1') > 0


def test_confusion_matrix_and_remote_aggregation():
    g = torch.Generator().manual_seed(1)
    pred, true = torch.randint(0, 4, (300,), generator=g), torch.randint(0, 4, (300,), generator=g)
    cm = ConfusionMatrix(num_classes=4)
    cm.add(pred, true)
    want = np.zeros((4, 4))
    for p, t in zip(pred.tolist(), true.tolist()):
        want[p, t] += 1
    assert np.array_equal(cm.matrix.numpy(), want)
    assert abs(cm.accuracy() - np.trace(want) / 300) < 1e-6
    other = ConfusionMatrix(num_classes=4)
    other.accumulate(cm); other.accumulate(cm)
    assert other.get()[0] == cm.get()[0]
    agg = ConfusionMatrix(num_classes=4)          # the reference breaks here
    agg.reduce_sites([cm.serialize(), cm.serialize()])
    assert agg.get() == cm.get()


def test_auc_matches_sklearn_with_ties():
    from sklearn import metrics as skm
    g = torch.Generator().manual_seed(2)
    prob = (torch.rand(400, generator=g) * 20).round() / 20      # many ties
    lab = (torch.rand(400, generator=g) < prob * 0.7 + 0.1).long()
    m = AUCROCMetrics()
    m.add(prob[:100], lab[:100]); m.add(prob[100:], lab[100:])
    assert abs(m.auc() - skm.roc_auc_score(lab.numpy(), prob.numpy())) < 1e-6
    assert len(m.probabilities) == 400
    r = AUCROCMetrics()
    r.reduce_sites([[0.7], [0.9]])
    assert r.get() == [0.8]


def test_dice_loss():
    o, t = torch.tensor([1., 1., 0., 0.]), torch.tensor([1., 0., 1., 0.])
    assert abs(float(dice_loss_binary(o, t)) - (1 - (2 * 1 + 1) / (2 + 2 + 1))) < 1e-6
    w = torch.tensor([0., 1., 1., 1.])
    assert torch.isfinite(dice_loss_binary(o, t, beta=2, weights=w))


def test_tensorutils_wire_format(tmp_path):
    net = torch.nn.Sequential(torch.nn.Linear(4, 4), torch.nn.Linear(4, 4))
    net(torch.randn(2, 4)).sum().backward()
    grads = tu.extract_grads(net, 'float16')
    path = tmp_path / 'grads.npy'
    tu.save_arrays(str(path), grads)           # all leading dims equal: np.array(list) would fail/mis-shape
    back = tu.load_arrays(str(path))
    assert back.dtype == object and back.shape == (4,) and back[0].dtype == np.float16 and back[0].shape == (4, 4)
    big, small = torch.zeros(1, 2, 10, 12), torch.ones(1, 3, 6, 8)
    assert tu.safe_concat(big, small).shape == (1, 5, 6, 8)
    assert tu.safe_concat(torch.zeros(1, 1, 8, 9, 10), torch.ones(1, 1, 4, 5, 6)).shape == (1, 2, 4, 5, 6)


def test_save_scores_and_cache(tmp_path):
    cache = {'log_header': 'Loss|Accuracy,F1', 'test_metrics': [[0.5, 0.9, 0.8]], 'obj': object(), 'n': {'t': torch.ones(1)}}
    save_scores(cache, str(tmp_path), file_keys=['test_metrics'])
    assert open(tmp_path / 'test_metrics.csv').read().split('\n')[:2] == ['Loss|Accuracy,F1', '0.5,0.9,0.8']
    save_cache(cache, str(tmp_path))
    assert json.load(open(tmp_path / 'logs.json'))['test_metrics'] == [[0.5, 0.9, 0.8]]


def test_arg_merge_priority():
    inp = {'task_id': 'tk', 'mode': 'train', 'batch_size': 3, 'agg_engine': 'powerSGD', 'num_folds': 2,
           'tk_args': {'batch_size': 5, 'epochs': 7}, 'powerSGD_args': {'epochs': 9, 'rank': 2},
           'tk_data_conf': {'epochs': 1, 'labels': 'l.csv'}}
    cache = {}
    COINNLocal(cache=cache, input=inp, state={}, learning_rate=0.5)
    assert cache['batch_size'] == 5           # <task>_args beats plain input
    assert cache['epochs'] == 9               # <engine>_args beats <task>_args
    assert cache['labels'] == 'l.csv' and cache['rank'] == 2
    assert cache['learning_rate'] == 0.5 and cache['validation_epochs'] == 1 and cache['patience'] == 31
    cache2 = {}
    with pytest.raises(AssertionError):
        COINNLocal(cache=cache2, input={'task_id': 'x', 'mode': 'train'}, state={})     # no split info


def test_checkpoint_layout_and_foreign_load(tmp_path):
    from coinstac_dinunet_b200.models import FSVTrainer
    from coinstac_dinunet_b200.distrib.nodes.remote import EmptyDataHandle
    cache = {'mode': 'train', 'seed': 1, 'learning_rate': 1e-3, 'num_class': 2}
    tr = FSVTrainer(data_handle=EmptyDataHandle(cache, {}, {}))
    tr.init_nn(init_model=True, init_optim=True, set_devices=True, init_weights=True)
    tr.nn['second'] = torch.nn.Linear(2, 2)
    tr.save_checkpoint(str(tmp_path / 'c.pt'))
    chk = torch.load(tmp_path / 'c.pt', weights_only=False)
    assert chk['source'] == 'coinstac' and set(chk['models']) == {'fs_net', 'second'} and set(chk['optimizers']) == {'adam'}
    before = tr.nn['fs_net'].classifier.weight.clone()
    with torch.no_grad():
        tr.nn['fs_net'].classifier.weight.zero_()
    tr.load_checkpoint(str(tmp_path / 'c.pt'))
    assert torch.equal(tr.nn['fs_net'].classifier.weight, before)
    torch.save(tr.nn['fs_net'].state_dict(), tmp_path / 'raw.pt')       # foreign checkpoint = bare state_dict
    tr.load_checkpoint(str(tmp_path / 'raw.pt'))


def test_imageutils_and_plotter(tmp_path):
    from coinstac_dinunet_b200.vision import imageutils as iu, plotter
    idx = list(iu.get_chunk_indexes((10, 10), (4, 4), (3, 3)))
    assert idx[0] == [0, 4, 0, 4] and idx[-1] == [6, 10, 6, 10] and len(idx) == 16   # last window repeats (as in the reference)
    patches = np.stack([np.full((4, 4), 7, np.uint8)] * len(idx))
    assert (iu.merge_patches(patches, (10, 10), (4, 4), (3, 3)) == 7).all()
    a, b = np.array([[255, 0], [255, 0]]), np.array([[255, 255], [0, 0]])
    assert iu.get_praf1(a, b) == {'Precision': 0.5, 'Recall': 0.5, 'Accuracy': 0.5, 'F1': 0.5}
    assert iu.get_rgb_scores(a, b)[0, 0].tolist() == [255, 255, 255]
    assert iu.get_pix_neigh(1, 1) == [(0, 1), (1, 2), (2, 1), (1, 0)] and len(iu.get_pix_neigh(1, 1, True)) == 8
    assert iu.get_chunk_indices_by_index((10, 10), (4, 4), [(0, 0), (9, 9)]) == [[0, 4, 0, 4], [6, 10, 6, 10]]
    assert iu.expand_and_mirror_patch((10, 10), [0, 4, 0, 4], (4, 4))[:4] == (0, 6, 0, 6)
    blob = np.zeros((8, 8), np.uint8); blob[1:3, 1:3] = 1; blob[5:8, 4:8] = 1
    assert iu.largest_cc(blob).sum() == 12
    cache = {'log_header': 'Loss|Accuracy,F1', 'train_log': [[1.0, .5, .4], [.8, .6, .5], [.6, .7, .6]]}
    plotter.plot_progress(cache, str(tmp_path), plot_keys=['train_log'])
    assert any(f.startswith('train_log_0') for f in os.listdir(tmp_path))


def test_bench_stdout_carries_only_the_json_line():
    """bench.py contract: ONE JSON line on stdout.  Library banners (NCCL prints its version on the first communicator)
    and stray prints must end up on stderr."""
    import json, os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import os, bench; bench._guard_stdout(); print('python noise'); os.system('echo c-level noise'); "
            "bench.emit({'metric': 'm', 'value': 1.5})")
    r = subprocess.run([sys.executable, '-c', code], cwd=root, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.splitlines()
    assert len(lines) == 1 and json.loads(lines[0]) == {'metric': 'm', 'value': 1.5}
    assert 'python noise' in r.stderr and 'c-level noise' in r.stderr


def test_native_library_exports_every_declared_symbol():
    """``ops/native.py`` declares ctypes signatures for the kernels' C entry points; after ``__graft_entry__.build()`` the
    in-tree ``_b200_ops.so`` must export all of them (it is loaded lazily, so a missing symbol would otherwise only
    surface on a GPU box)."""
    import ctypes, os, re
    import pytest
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    so = os.path.join(here, 'coinstac_dinunet_b200', 'ops', '_b200_ops.so')
    if not os.path.exists(so):
        pytest.skip('kernel library not built (run __graft_entry__.build())')
    try:
        lib = ctypes.CDLL(so)
    except OSError as exc:                                   # no CUDA runtime on this machine
        pytest.skip(f'cannot load {so}: {exc}')
    src = open(os.path.join(here, 'coinstac_dinunet_b200', 'ops', 'native.py')).read()
    names = sorted(set(re.findall(r'd\.(coinn_\w+)\.', src)))
    assert len(names) >= 25
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_nativize_plans_user_models_without_touching_state_dict():
    """ops.nativize on a user-defined CNN: the plan covers conv stacks (with channel padding) and Linear(+BN1d)(+ReLU)
    groups, module names / state_dict keys / parameters are untouched and the CPU forward is the original one."""
    import torch
    from torch import nn
    from coinstac_dinunet_b200.ops.nativize import NativeSequential, _padded_pair, nativize

    class UserNet(nn.Module):
        def __init__(self):
            super().__init__()
            self.features = nn.Sequential(
                nn.Conv3d(1, 8, 3, padding=1), nn.BatchNorm3d(8), nn.ReLU(), nn.MaxPool3d(2),
                nn.Conv3d(8, 24, 3, padding=1, bias=False), nn.BatchNorm3d(24), nn.ReLU(inplace=True), nn.MaxPool3d(2),
                nn.Conv3d(24, 48, 3, padding=1), nn.BatchNorm3d(48), nn.ReLU(), nn.MaxPool3d(2),
                nn.Dropout3d(0.0))
            self.head = nn.Sequential(nn.Flatten(), nn.Linear(48 * 8, 64), nn.BatchNorm1d(64), nn.ReLU(), nn.Linear(64, 32),
                                      nn.ReLU(), nn.Linear(32, 2))

        def forward(self, x):
            return self.head(self.features(x))

    torch.manual_seed(0)
    net = UserNet()
    keys = list(net.state_dict().keys())
    params = [id(p) for p in net.parameters()]
    x = torch.randn(2, 1, 16, 16, 16)
    net.eval()
    want = net(x)
    report = []
    out = nativize(net, report=report)
    assert out is net and isinstance(net.features, NativeSequential) and isinstance(net.head, NativeSequential)
    assert list(net.state_dict().keys()) == keys and [id(p) for p in net.parameters()] == params
    assert torch.equal(net(x), want)                                   # CPU: original path
    kinds = [(r[0], r[1]) for r in report]
    assert ('features', 'conv_stack') in kinds and ('head', 'linear_bn_relu') in kinds
    assert ('head', 'linear_relu') in kinds and ('head', 'linear') in kinds
    stack = [r for r in report if r[1] == 'conv_stack'][0][2]
    assert stack == [(1, 8), (8, 24), (24, 48)]
    assert _padded_pair(1, 8) == (16, 32) and _padded_pair(24, 48) == (32, 64) and _padded_pair(48, 96) == (64, 128)
    assert _padded_pair(300, 10) is None
    plan = net.features._native_plan
    assert plan[0][0] == 'stack' and plan[-1][0] == 'mod'              # Dropout3d stays a torch module
    # strides / kernel sizes the kernels do not implement stay on the torch path
    other = nn.Sequential(nn.Conv3d(4, 8, 5, padding=2), nn.BatchNorm3d(8), nn.ReLU(), nn.MaxPool3d(2))
    assert not isinstance(nativize(other), NativeSequential)


def test_affinity_helpers_parse_topology_and_never_raise():
    from coinstac_dinunet_b200.utils import affinity
    assert affinity._parse_cpulist('0-3,8,10-11\n') == {0, 1, 2, 3, 8, 10, 11}
    assert affinity._parse_cpulist('') == set()
    rep = affinity.pin_to_gpu(0)            # no GPU here: must report "not pinned" instead of failing
    assert rep['gpu'] == 0 and rep['pinned'] in (False, True)


def test_custom_example_model_is_nativizable():
    import importlib.util, os
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location('custom_local', os.path.join(here, 'examples', 'custom', 'local.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    from coinstac_dinunet_b200.ops import nativize
    report = []
    net = nativize(mod.MyNet((16, 16, 16)), report=report)
    kinds = {r[1] for r in report}
    assert {'conv_stack', 'linear_bn_relu', 'linear'} <= kinds
    import torch
    net.eval()
    assert net(torch.randn(2, 1, 16, 16, 16)).shape == (2, 2)


def test_coefficient_space_power_iteration_is_the_truncated_svd():
    """The algorithm of csrc/lowrank.cu::lowrank_eig_kernel (PyTorch twin): top-k triplets of B C^T from the two n x n Gram
    matrices only - reconstruction within 2 % of the optimal rank-k error, right factors orthonormal."""
    import torch
    from coinstac_dinunet_b200.ops.lowrank import lowrank_factor_reference
    torch.manual_seed(0)
    for rows_b, rows_c, n, rank in ((256, 67, 16, 10), (32, 33, 8, 10), (64, 300, 32, 4), (2, 33, 16, 10)):
        decay = torch.logspace(0, -3, n)
        B, C = torch.randn(rows_b, n) * decay, torch.randn(rows_c, n)
        left, right = lowrank_factor_reference(B, C, rank, 40, 1e-6)
        k = left.shape[1]
        assert k == min(rank, rows_b, rows_c, n)
        full = B.double() @ C.double().t()
        U, S, Vh = torch.linalg.svd(full, full_matrices=False)
        best = (U[:, :k] * S[:k]) @ Vh[:k]
        err, opt = float((full - left.double() @ right.double().t()).norm()), float((full - best).norm())
        assert err <= opt * 1.02 + 1e-6 * float(full.norm()), (rows_b, rows_c, n, err, opt)
        gram = right.t() @ right
        keep = gram.diagonal() > 0.5
        assert torch.allclose(gram[keep][:, keep], torch.eye(int(keep.sum())), atol=5e-2)


def test_powersgd_plan_tables_and_host_helpers():
    """Descriptor tables of the batched PowerSGD kernels (built on the host, CPU is fine): every matrix gets disjoint P / Q
    ranges, small or 1-D tensors are exchanged uncompressed, the gather / scatter segments are inverse to each other."""
    import os
    import numpy as np
    import pytest
    import torch
    so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'coinstac_dinunet_b200', 'ops', '_b200_ops.so')
    if not os.path.exists(so):
        pytest.skip('kernel library not built')
    try:
        from coinstac_dinunet_b200.ops.lowrank import PowerSGDPlan
        params = [torch.nn.Parameter(torch.zeros(*s)) for s in ((256, 66), (256,), (128, 256), (2, 32), (64, 16, 3, 3, 3), (64,))]
        offsets, off = [], 0
        for p in params:
            offsets.append(off); off += (p.numel() + 7) // 8 * 8
        plan = PowerSGDPlan(params, offsets, 2, 'cpu')
    except OSError as exc:
        pytest.skip(f'cannot load the kernel library here: {exc}')
    assert [m[:3] for m in plan.mats] == [(0, 256, 66), (2, 128, 256), (4, 64, 432)]          # (2, 32) has <= rank rows: dense
    assert [l[0] for l in plan.low] == [1, 3, 5] and plan.low_numel == 256 + 64 + 64
    assert plan.p_numel == 2 * (256 + 128 + 64) and plan.q_numel == 2 * (66 + 256 + 432)
    desc = plan.desc.numpy().view([('g', '<i8'), ('p', '<i8'), ('q', '<i8'), ('n', '<i4'), ('m', '<i4')])
    assert list(desc['g']) == [offsets[0], offsets[2], offsets[4]] and list(desc['p']) == [0, 512, 768]
    g, s = plan.seg_gather.numpy().view('<i8').reshape(-1, 3), plan.seg_scatter.numpy().view('<i8').reshape(-1, 3)
    assert np.array_equal(g[:, 0], s[:, 1]) and np.array_equal(g[:, 1], s[:, 0]) and np.array_equal(g[:, 2], s[:, 2])
    assert g[0, 1] == plan.q_numel                        # rank-1 gradients ride behind the Q factors


def test_lagged_readback_and_running_scores_on_cpu():
    import torch
    from coinstac_dinunet_b200.parallel.nvlink_learner import _LaggedReadback, _RunningScores
    rb = _LaggedReadback(torch.device('cpu'), slots=4, lag=2)
    for i in range(7):
        rb.push(torch.tensor(float(i)))
    assert rb.finish() == 6.0 and rb.count == 7
    cache = {}
    assert _LaggedReadback.for_site(cache, torch.device('cpu')) is _LaggedReadback.for_site(cache, torch.device('cpu'))

    class T:
        def new_averages(self):
            from coinstac_dinunet_b200.metrics import COINNAverages
            return COINNAverages(1)

        def new_metrics(self):
            from coinstac_dinunet_b200.metrics import Prf1a
            return Prf1a()

        def reduce_iteration(self, its):
            return {k: v for k, v in its[0].items()}
    t = T()
    sc = _RunningScores(t)
    for i in range(3):
        a, m = t.new_averages(), t.new_metrics()
        a.add(float(i), 4)
        m.add(torch.tensor([1, 0, 1, 1]), torch.tensor([1, 0, 0, 1]))
        sc.add([{'averages': a, 'metrics': m, 'loss': torch.tensor(float(i))}])
    out = sc.result()
    assert abs(out['averages'].get()[0] - 1.0) < 1e-6 and float(out['loss']) == 2.0
    assert out['metrics'].get()[0] == 0.75                # accuracy over the 12 accumulated predictions


def test_arena_bucket_layout_and_shard_ownership_math():
    """Host-side bookkeeping of the fused data plane, exercised without a GPU on a bare DistArena object: buckets are formed
    in backward order over whole parameters and tile the arena; the (rank, lo, hi) ownership ranges of two-shot / NVLS launch
    units follow the kernel's ceil(nvec / S) vectors-per-rank rule and tile every sharded unit exactly once."""
    import torch
    from coinstac_dinunet_b200.parallel import arena as A
    ar = object.__new__(A.DistArena)
    sizes = [432, 16, 16, 13824, 32, 32, 55296, 64, 64, 2359296, 256, 16384, 64, 128, 2]
    ar.params = [torch.nn.Parameter(torch.zeros(n)) for n in sizes]
    ar.offsets, off = [], 0
    for n in sizes:
        ar.offsets.append(off); off += A._round_up(n, A._ALIGN)
    ar.world, ar.rank, ar.backend, ar.variant, ar.device, ar.group = 8, 3, 'nvlink', 'auto', torch.device('cpu'), None
    ar.numel = A._round_up(off, 4 * ar.world)

    class _Buf:
        multicast_ptr = 1
    ar.wire_buf, ar.grad_buf, ar.param_buf = None, _Buf(), _Buf()
    # --- bucket construction (the part of enable_overlap that needs no device)
    cap = (1 << 20) // 4
    buckets, stop, members = [], ar.numel, []
    for i in range(len(ar.params) - 1, -1, -1):
        members.append(i)
        start = ar.offsets[i]
        if stop - start >= cap or i == 0:
            start = 0 if i == 0 else start
            buckets.append({'offset': start, 'numel': stop - start, 'params': members})
            stop, members = start, []
    assert sum(b['numel'] for b in buckets) == ar.numel and buckets[-1]['offset'] == 0
    assert all(b['offset'] % 4 == 0 and b['numel'] % 4 == 0 for b in buckets)
    assert buckets[0]['params'][-1] == 9                       # the 9.4 MB FC1 weight closes the first (head) bucket
    ar._overlap = {'buckets': buckets}
    # --- ownership ranges
    ranges = ar.owner_ranges()
    for b in buckets:
        variant = ar._pick_variant(b['numel'] * 4)
        mine = sorted((lo, hi, q) for q, lo, hi in ranges if b['offset'] <= lo < b['offset'] + b['numel'])
        if variant == 'one_shot':
            assert not mine                                    # replicated update: nothing to gather
            continue
        assert variant == 'nvls'
        assert mine[0][0] == b['offset'] and mine[-1][1] == b['offset'] + b['numel']
        assert all(a[1] == c[0] for a, c in zip(mine, mine[1:]))         # contiguous, no overlap
        shard = -(-(b['numel'] // 4) // ar.world) * 4
        assert all(hi - lo <= shard for lo, hi, _ in mine) and [q for _, _, q in mine] == sorted(q for _, _, q in mine)
    ar.world = 2                                               # at two sites NVLS brings nothing: two-shot
    assert ar._pick_variant(16 << 20) == 'two_shot' and ar._pick_variant(1024) == 'one_shot'


def test_bench_helpers():
    """bench.py bookkeeping that does not need a GPU: R1 lookup, clock-record merge, host gap statistics, CLI defaults."""
    import importlib.util, os, sys
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location('bench_mod', os.path.join(here, 'bench.py'))
    bench = importlib.util.module_from_spec(spec)
    argv, sys.argv = sys.argv, ['bench.py']
    try:
        spec.loader.exec_module(bench)
        a = bench.parse()
    finally:
        sys.argv = argv
    assert (a.gpus, a.impl, a.model, a.dtype, a.input_dtype, a.overlap) == (1, 'ours', 'vbm', 'bf16', 'fp32', -1)
    assert a.warmup >= 3 and a.steps > 0
    assert bench.r1_number('vbm', 1) and bench.r1_number('vbm', 8) and bench.r1_number('vbm', 3) is None
    assert bench.gap_stats([0.0, 0.002, 0.004, 0.0075]) == {'p50': 2.0, 'max': 3.5} and bench.gap_stats([0.0]) is None
    m = bench.merge_clocks({'sm_mhz': 1965.0, 'sm_max_mhz': 1965.0, 'reasons': [], 'samples': 10, 'power_w_max': 300.0},
                           {'sm_mhz': 1950.0, 'sm_max_mhz': 1965.0, 'reasons': ['sw_power_cap'], 'samples': 5}, None)
    assert m['samples'] == 15 and m['reasons'] == ['sw_power_cap'] and m['sm_max_mhz'] == 1965.0 and m['power_w_max'] == 300.0
    assert 'VBM' in bench.metric_name('vbm') and 'FreeSurfer' in bench.metric_name('fs')


class _TorchConvBlock:
    """fp32 PyTorch stand-in for ``ops.vbm.ConvBnReluPoolFn`` (same calling convention, channels-last activations) so the
    host-side algebra of ``ops.nativize._conv_stack`` - channel padding, conv-bias folding, running-statistics handling - can be
    checked on CPU against the user's original modules."""

    @staticmethod
    def apply(h, w, g, b, rm, rv, eps, mom, training, backend, nbt=None):
        F = torch.nn.functional
        x = h.unsqueeze(1) if h.dim() == 4 else h.permute(0, 4, 1, 2, 3)
        y = F.conv3d(x.float(), w.float(), padding=1)
        rm2, rv2 = rm.clone(), rv.clone()                    # (autograd keeps the buffers it is given; the kernels do not)
        y = F.batch_norm(y, rm2, rv2, g, b, training, mom, eps)
        with torch.no_grad():
            rm.copy_(rm2), rv.copy_(rv2)
        if nbt is not None:
            nbt += 1
        return F.max_pool3d(F.relu(y), 2).permute(0, 2, 3, 4, 1).contiguous()


@pytest.mark.parametrize('chans,bias', [((1, 16, 32), False), ((3, 20, 40, 50), True), ((16, 32, 24), True), ((5, 64), False)])
def test_nativize_conv_stack_algebra_matches_the_original_modules(chans, bias, monkeypatch):
    """Zero-padded channels stay exactly zero through BN(gamma 1, beta 0)+ReLU+pool, a conv bias in front of BatchNorm only
    shifts the tracked mean, and the user's BatchNorm buffers end up identical to those of the unmodified module - in
    training (outputs, input / weight / BN gradients, running stats, num_batches_tracked) and in eval mode."""
    import copy
    from torch import nn
    import importlib
    from coinstac_dinunet_b200.ops import vbm
    nz = importlib.import_module('coinstac_dinunet_b200.ops.nativize')   # (the package re-exports the function under this name)
    monkeypatch.setattr(vbm, 'ConvBnReluPoolFn', _TorchConvBlock)
    torch.manual_seed(3)
    layers = []
    for ci, co in zip(chans, chans[1:]):
        layers += [nn.Conv3d(ci, co, 3, padding=1, bias=bias), nn.BatchNorm3d(co), nn.ReLU(), nn.MaxPool3d(2)]
    ref = nn.Sequential(*layers).double()
    for m in ref:
        if isinstance(m, nn.BatchNorm3d):
            m.weight.data.uniform_(0.5, 1.5), m.bias.data.normal_(0, 0.2)
            m.running_mean.normal_(0, 0.1), m.running_var.uniform_(0.5, 2.0)
    ref = ref.float()
    ours = copy.deepcopy(ref)
    plan = nz._plan(list(ours.children()))
    assert [st[0] for st in plan] == ['stack'] and len(plan[0][1]) == len(chans) - 1
    side = 2 ** (len(chans) - 1) * 2
    # bf16-representable inputs: _conv_stack hands the kernels bf16 activations, the comparison should see only the algebra
    x = torch.randn(2, chans[0], side, side, side).bfloat16().float()

    for training in (True, False):
        ref.train(training), ours.train(training)
        xr, xo = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
        yr = ref(xr)
        # the stand-in keeps fp32 between blocks; the real kernels round to bf16 there, which is what the GPU tests bound
        yo = nz._conv_stack(plan[0][1], xo if chans[0] != 1 else xo[:, 0], training)
        assert yo.shape == yr.shape
        assert torch.allclose(yo, yr, rtol=1e-4, atol=1e-4), float((yo - yr).abs().max())
        if training:
            (yr.square().mean()).backward()
            (yo.square().mean()).backward()
            for (n, pr), (_, po) in zip(ref.named_parameters(), ours.named_parameters()):
                if n.endswith('.bias') and isinstance(dict(ref.named_modules())[n.rsplit('.', 1)[0]], nn.Conv3d):
                    assert po.grad is None or float(po.grad.abs().max()) < 1e-4        # a bias before BN has zero gradient
                    continue
                scale = float(pr.grad.abs().max()) + 1e-12
                assert float((po.grad - pr.grad).abs().max()) <= 1e-3 * scale + 1e-7, n
        for (n, br), (_, bo) in zip(ref.named_buffers(), ours.named_buffers()):
            assert torch.allclose(bo.float(), br.float(), rtol=1e-4, atol=1e-5), n


class _FlakyDataset:
    """Module-level (picklable for worker processes): every third sample 'fails to load' (returns None); the others carry a
    NumPy random number so worker seeding is observable."""

    def __init__(self, n=12):
        self.n = n

    def __len__(self):
        return self.n

    def __getitem__(self, ix):
        import numpy as _np
        if ix % 3 == 2:
            return None
        return {'ix': torch.tensor(ix), 'noise': torch.tensor(float(_np.random.rand()))}


def test_safe_collate_drops_failed_samples_and_workers_are_seeded():
    """``safe_collate`` silently drops samples that failed to load (ref data.py:23-27); with ``seed_all`` every worker process
    seeds NumPy from torch's per-worker seed, so augmentation noise is reproducible for a fixed base seed and differs between
    workers (ref data.py:96-99)."""
    from coinstac_dinunet_b200.data import COINNDataHandle, safe_collate
    batch = safe_collate([{'x': torch.ones(2)}, None, {'x': torch.zeros(2)}, {}])
    assert batch['x'].shape == (2, 2)

    def run(seed):
        dh = COINNDataHandle(cache={'seed_all': True, 'num_workers': 2, 'batch_size': 3, 'seed': 5}, input={}, state={})
        torch.manual_seed(seed)
        loader = dh.get_loader('train', dataset=_FlakyDataset(), shuffle=False)
        assert loader.worker_init_fn is not None and loader.num_workers == 2
        rows = [(b['ix'].tolist(), b['noise'].tolist()) for b in loader]
        return rows
    a, b, c = run(11), run(11), run(12)
    assert [ix for ix, _ in a] == [[0, 1], [3, 4], [6, 7], [9, 10]]          # index 2, 5, 8, 11 were dropped, no crash
    assert a == b and [n for _, n in a] != [n for _, n in c]                  # reproducible per base seed
    assert a[0][1] != a[1][1]                                                 # worker 0 and worker 1 draw different streams


def test_shipped_library_is_blackwell_native():
    """The in-tree ``_b200_ops.so`` must be an sm_100a build whose SASS really uses the 5th-generation tensor cores, TMEM, TMA
    and multimem - not a CUDA-core fallback that happens to export the same symbols (B200_PROFILING.md names the mnemonics)."""
    import re
    import shutil
    import subprocess
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    so = os.path.join(here, 'coinstac_dinunet_b200', 'ops', '_b200_ops.so')
    tool = shutil.which('cuobjdump') or '/usr/local/cuda/bin/cuobjdump'
    if not os.path.exists(so) or not os.path.exists(tool):
        pytest.skip('needs the built library and cuobjdump')
    elf = subprocess.run([tool, '-lelf', so], capture_output=True, text=True, timeout=120).stdout
    archs = set(re.findall(r'sm_\d+a?', elf))
    assert archs == {'sm_100a'}, archs
    sass = subprocess.run([tool, '-sass', so], capture_output=True, text=True, timeout=600).stdout
    want = {'UTCHMMA': 'tcgen05.mma (bf16)', 'UTCQMMA': 'tcgen05.mma block-scaled (MX-FP8)', 'UTMALDG': 'TMA tensor load',
            'LDTM': 'tcgen05.ld (TMEM -> registers)', 'UTCCP': 'tcgen05.cp (scale factors -> TMEM)',
            'LDGMC': 'multimem.ld_reduce (NVLS)', 'UTCBAR': 'tcgen05.commit -> mbarrier'}
    counts = {k: len(re.findall(r'\b' + k + r'\b', sass)) for k in want}
    assert all(counts.values()), {want[k]: v for k, v in counts.items()}
    assert counts['UTCHMMA'] > 500 and counts['UTMALDG'] > 1000
