"""Secondary protocol flows on CPU: SiteRunner, pre-training + weight broadcast, test-only mode, sparse test
datasets, the in-memory dataset / pinned collate, user-supplied learner / reducer classes."""
import json
import os

import numpy as np
import pytest
import torch

from coinstac_dinunet_b200 import COINNLearner, COINNReducer, SiteRunner
from coinstac_dinunet_b200.data import COINNDataHandle
from coinstac_dinunet_b200.models import FSVDataset, FSVTrainer, InMemorySynthetic, write_synthetic_site
from coinstac_dinunet_b200.models.common import pinned_collate


def test_site_runner_trains_one_site_offline(tmp_path):
    data = tmp_path / 'sim'
    base = data / 'input' / 'local0' / 'simulatorRun'
    os.makedirs(base)
    write_synthetic_site(str(base), 40, (66,), seed=0)
    spec = [{k: {'value': v} for k, v in dict(mode='train', data_dir='data', labels_file='labels.json', num_class=2,
                                              batch_size=4, epochs=3, split_ratio=[0.6, 0.2, 0.2], learning_rate=1e-2,
                                              monitor_metric='f1', metric_direction='maximize', patience=5).items()}]
    with open(data / 'inputspec.json', 'w') as fp:
        json.dump(spec, fp)
    runner = SiteRunner('fsv', data_path=str(data), site_index=0)
    out = runner.run(FSVTrainer, FSVDataset, COINNDataHandle)
    assert out['phase'] == 'pre_computation'
    odir = runner.state['outputDirectory']
    assert os.path.exists(os.path.join(odir, 'fsv', 'splits', 'SPLIT.json'))
    assert os.path.exists(os.path.join(odir, 'fsv', 'fold_0', 'logs.json'))
    assert len(runner.cache['train_log']) > 0 and len(runner.cache['validation_log']) == 3
    assert os.path.exists(os.path.join(odir, 'weights.tar'))          # best pre-training weights in the transfer dir


def test_pretraining_broadcasts_weights(fs_sites):
    eng = fs_sites(spec={'num_folds': None, 'split_ratio': [0.6, 0.2, 0.2], 'epochs': 1})
    eng.run_nodes(FSVTrainer, FSVDataset, local_kw={'pretrain_args': {'epochs': 2}}, max_rounds=1000)
    phases = [t['remote'] for t in eng.trace]
    assert 'pre_computation' in phases and phases[-2] == 'success'
    # only the site with most training data pre-trains; both then start from its broadcast weights
    assert os.path.exists(os.path.join(eng.site_state['local1']['baseDirectory'], 'pretrained_weights.tar'))
    a, b = (torch.cat([p.detach().reshape(-1) for p in eng.site_cache[s]['nn']['fs_net'].parameters()]) for s in eng.site_ids)
    assert torch.equal(a, b)


def test_sparse_test_datasets_and_save_predictions(fs_sites):
    seen = []

    class T(FSVTrainer):
        def save_predictions(self, dataset, its):
            seen.append((len(dataset), type(its['prediction']).__name__))
            return {}
    eng = fs_sites(spec={'num_folds': None, 'split_ratio': [0.6, 0.2, 0.2], 'epochs': 1, 'load_sparse': True})
    eng.run_nodes(T, FSVDataset, max_rounds=1000)
    assert seen and all(n == 1 for n, _ in seen)                  # one dataset per test subject
    assert all(kind == '_LazyCollect' for _, kind in seen)        # reduce_iteration hands lazily-collected outputs


def test_custom_learner_and_reducer_classes(fs_sites):
    calls = {'learner': 0, 'reducer': 0}

    class MyLearner(COINNLearner):
        def to_reduce(self):
            calls['learner'] += 1
            return super().to_reduce()

    class MyReducer(COINNReducer):
        def reduce(self):
            calls['reducer'] += 1
            return super().reduce()
    eng = fs_sites(spec={'num_folds': None, 'split_ratio': [0.6, 0.2, 0.2], 'epochs': 1, 'agg_engine': 'custom'})
    eng.run_nodes(FSVTrainer, FSVDataset, learner_cls=MyLearner, reducer_cls=MyReducer, max_rounds=1000)
    assert calls['learner'] > 0 and calls['reducer'] > 0 and eng.trace[-2]['remote'] == 'success'


def test_in_memory_dataset_and_pinned_collate():
    ds = InMemorySynthetic(shape=(3, 4), num_class=2, seed=1, cache={'synthetic_distinct': 8})
    ds.add([str(i) for i in range(20)])
    a, b = ds[3], ds[11]                                          # indexes the 8-sample pool modulo
    assert torch.equal(a['inputs'], b['inputs']) and a['inputs'].shape == (3, 4)
    batch = pinned_collate([ds[i] for i in range(4, 8)])
    assert batch['inputs'].shape == (4, 3, 4) and batch['inputs']._base is ds._x    # zero-copy slice
    mixed = pinned_collate([ds[0], ds[2], ds[5]])
    assert mixed['inputs'].shape == (3, 3, 4) and torch.equal(mixed['inputs'][1], ds[2]['inputs'])


def test_test_only_mode_uses_pretrained_checkpoint(fs_sites, tmp_path):
    train = fs_sites(spec={'num_folds': None, 'split_ratio': [0.6, 0.2, 0.2], 'epochs': 1})
    train.run_nodes(FSVTrainer, FSVDataset, max_rounds=1000)
    ckpt = os.path.join(train.site_state['local0']['outputDirectory'], 'fsv', 'fold_0', 'latest.fsv-0.pt')
    assert os.path.exists(ckpt)
    chk = torch.load(ckpt, weights_only=False)
    assert chk['source'] == 'coinstac' and 'fs_net' in chk['models'] and 'adam' in chk['optimizers']


def test_example_computations_run_in_the_simulator():
    import importlib.util, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location('run_simulator', os.path.join(root, 'examples', 'run_simulator.py'))
    sim = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sim)
    for which in ('fsv', 'vbm', 'custom'):                   # (custom: local.py / remote.py import each other as siblings)
        eng = sim.main(which)
        assert eng.trace[-2]['remote'] == 'success', which
    assert json.load(open(os.path.join(root, 'examples', 'vbm', 'compspec.json')))['computation']['input']['transport']['default'] == 'nvlink'


def test_gradient_accumulation_and_early_stop(fs_sites):
    """``local_iterations=2`` (micro-batches accumulate before every reduce, reference learner.py:32-47) and
    ``patience=1`` (remote early stop, remote.py:286-287) through the whole protocol."""
    eng = fs_sites(spec={'num_folds': None, 'split_ratio': [0.6, 0.2, 0.2], 'epochs': 30, 'patience': 1,
                         'local_iterations': 2, 'learning_rate': 1e-6})          # lr ~ 0: validation score cannot improve
    rounds = eng.run_nodes(FSVTrainer, FSVDataset, max_rounds=3000)
    assert eng.trace[-2]['remote'] == 'success'
    # 30 epochs of 15 / 11 samples at batch 4 x 2 micro-batches would need > 100 reduce rounds; patience stops long before
    assert rounds < 60, rounds
    a, b = (torch.cat([p.detach().reshape(-1) for p in eng.site_cache[s]['nn']['fs_net'].parameters()]) for s in eng.site_ids)
    assert torch.equal(a, b)


def test_multiclass_protocol_uses_confusion_matrix(tmp_path):
    """num_class = 3 -> ConfusionMatrix metrics on the sites and their (fixed) aggregation on the remote."""
    import csv
    from coinstac_dinunet_b200.engine import InProcessEngine
    from coinstac_dinunet_b200.models import write_synthetic_site
    spec = dict(task_id='fsv', mode='train', data_dir='data', labels_file='labels.json', input_size=66, num_class=3,
                batch_size=4, epochs=2, num_folds=None, split_ratio=[0.6, 0.2, 0.2], learning_rate=1e-2, seed=3,
                monitor_metric='f1', metric_direction='maximize', log_header='Loss|Accuracy,F1', verbose=False)
    eng = InProcessEngine(tmp_path / 'work', n_sites=2, inputspec=spec)
    for i, site in enumerate(eng.site_ids):
        write_synthetic_site(eng.site_state[site]['baseDirectory'], 30, (66,), num_class=3, seed=i)
    eng.run_nodes(FSVTrainer, FSVDataset, max_rounds=1000)
    assert eng.trace[-2]['remote'] == 'success'
    from coinstac_dinunet_b200.metrics import ConfusionMatrix
    from coinstac_dinunet_b200.distrib.nodes.remote import EmptyDataHandle
    tr = FSVTrainer(data_handle=EmptyDataHandle(eng.site_cache['local0'], {}, eng.site_state['local0']))
    assert isinstance(tr.new_metrics(), ConfusionMatrix)
    path = os.path.join(eng.remote_state['outputDirectory'], 'fsv', 'global_test_metrics.csv')
    rows = list(csv.reader(open(path)))
    assert len(rows) >= 2 and all(0.0 <= float(v) <= 1.0 for v in rows[-1][1:] if v not in ('', 'nan'))


def test_user_defined_metric_travels_through_the_protocol(fs_sites):
    """README highlight 5 of the reference: a ``COINNMetrics`` subclass written by the user (here: a Brier score, lower is
    better) is created through ``new_metrics``, serialised by the sites, reduced by the aggregator and drives model selection
    (``monitor_metric='brier'``, ``metric_direction='minimize'``) - ref metrics.py:17-84, remote.py:105-141, 281-284."""
    from coinstac_dinunet_b200.metrics import COINNMetrics

    class Brier(COINNMetrics):
        def __init__(self, **kw):
            super().__init__(**kw)
            self.sq, self.n, self._reduced = 0.0, 0, None

        def add(self, prob, true):
            self.sq += float(((prob.detach().float() - true.float()) ** 2).sum())
            self.n += int(true.numel())

        def accumulate(self, other):
            self.sq, self.n = self.sq + other.sq, self.n + other.n

        def reset(self):
            self.sq, self.n = 0.0, 0

        @property
        def brier(self):
            return self._reduced if self._reduced is not None else self.sq / max(self.n, 1)

        def get(self):
            return [round(self.brier, 5)]

        def serialize(self, **kw):
            return [self.sq, self.n]

        def reduce_sites(self, scores):
            tot = np.asarray(scores, dtype=np.float64).sum(0) if len(scores) else np.zeros(2)
            self._reduced = float(tot[0] / max(tot[1], 1))

    class BrierTrainer(FSVTrainer):
        def new_metrics(self):
            return Brier()

        def iteration(self, batch):
            x, y = self._inputs(batch)
            logits = self.nn['fs_net'](x)
            loss = torch.nn.functional.cross_entropy(logits, y)
            avg, met = self.new_averages(), self.new_metrics()
            avg.add(loss.detach(), len(y))
            met.add(torch.softmax(logits, 1)[:, 1], y)
            return {'loss': loss, 'averages': avg, 'metrics': met}

    eng = fs_sites(spec={'num_folds': None, 'split_ratio': [0.6, 0.2, 0.2], 'epochs': 4, 'monitor_metric': 'brier',
                         'metric_direction': 'minimize', 'log_header': 'Loss|Brier'})
    eng.run_nodes(BrierTrainer, FSVDataset, max_rounds=3000)
    rc = eng.remote_cache
    assert eng.trace[-2]['remote'] == 'success'
    val = [row[1] for row in rc['validation_log']]
    assert len(val) >= 4 and all(0.0 <= v <= 1.0 for v in val)
    assert rc['best_val_score'] == pytest.approx(min(val), abs=2e-4)         # minimised (improvements below score_delta do not count)
    train = [row[1] for row in rc['train_log']]
    assert train[-1] < train[0] and train[0] > 0.01                           # and it learns
    assert len(rc['test_metrics'][0]) == 2                                    # [loss, brier]
    out_dir = os.path.join(eng.remote_state['outputDirectory'], 'fsv')
    assert os.path.exists(os.path.join(out_dir, 'global_test_metrics.csv'))


def test_profile_flag_reports_per_site_phase_timings(tmp_path):
    """README highlight 7 of the reference ("realtime profiling each site by specifying in compspec"): ``profile: true`` makes
    every site time its forward/backward and reduce/update phases and ship the table with its output."""
    from coinstac_dinunet_b200.engine import InProcessEngine
    spec = dict(task_id='fsv', mode='train', data_dir='data', labels_file='labels.json', input_size=66, num_class=2,
                batch_size=4, epochs=1, num_folds=None, split_ratio=[0.6, 0.2, 0.2], learning_rate=1e-2, seed=7,
                transport='nvlink', profile=True, verbose=False)
    eng = InProcessEngine(tmp_path / 'work', n_sites=1, inputspec=spec)
    write_synthetic_site(eng.site_state['local0']['baseDirectory'], 24, (66,), seed=0)
    eng.run_nodes(FSVTrainer, FSVDataset, max_rounds=500)
    log = eng.site_cache['local0'].get('profile_log')
    assert log, 'no profile table in the site cache'
    last = log[-1] if isinstance(log, list) else log
    text = json.dumps(last, default=str)
    assert 'forward_backward' in text and 'reduce_update' in text


def test_multi_network_training_scheme_with_a_custom_learner(fs_sites):
    """README highlights 1 + 6 of the reference: a trainer that owns TWO networks with their own optimizers (encoder + head),
    and a user learner that puts both on the wire.  With the stock learner only the first network is exchanged (the
    reference's rule, learner.py:30-37); with the custom one every replica of both networks stays identical, and the
    checkpoint keeps both models and both optimizers (the reference drops all but the last)."""
    from torch import nn

    class TwoNetTrainer(FSVTrainer):
        def _init_nn_model(self):
            self.nn['encoder'] = nn.Sequential(nn.Linear(66, 16), nn.ReLU())
            self.nn['head'] = nn.Linear(16, 2)

        def _init_optimizer(self):
            self.optimizer['enc_opt'] = torch.optim.Adam(self.nn['encoder'].parameters(), lr=1e-2)
            self.optimizer['head_opt'] = torch.optim.SGD(self.nn['head'].parameters(), lr=5e-2)

        def iteration(self, batch):
            x, y = self._inputs(batch)
            logits = self.nn['head'](self.nn['encoder'](x))
            loss = torch.nn.functional.cross_entropy(logits, y)
            avg, met = self.new_averages(), self.new_metrics()
            avg.add(loss.detach(), len(y))
            met.add(logits.argmax(1), y)
            return {'loss': loss, 'averages': avg, 'metrics': met}

    class AllNetsLearner(COINNLearner):
        @property
        def model(self):                                   # every network, in the trainer's order, as one parameter list
            return nn.ModuleList(list(self.trainer.nn.values()))

        def backward(self):
            for opt in self.trainer.optimizer.values():
                opt.zero_grad()
            return super().backward()

        def step(self):
            out = super().step()                           # installs the averaged gradients, steps the first optimizer
            for opt in list(self.trainer.optimizer.values())[1:]:
                opt.step()
            return out

    def flat(cache, name):
        return torch.cat([p.detach().reshape(-1) for p in cache['nn'][name].parameters()])

    spec = {'num_folds': None, 'split_ratio': [0.6, 0.2, 0.2], 'epochs': 2, 'agg_engine': 'custom'}
    eng = fs_sites(spec=spec)
    eng.run_nodes(TwoNetTrainer, FSVDataset, learner_cls=AllNetsLearner, max_rounds=2000)
    assert eng.trace[-2]['remote'] == 'success'
    a, b = eng.site_cache['local0'], eng.site_cache['local1']
    assert torch.equal(flat(a, 'encoder'), flat(b, 'encoder')) and torch.equal(flat(a, 'head'), flat(b, 'head'))
    chk = torch.load(os.path.join(a['log_dir'], a['latest_nn_state']), weights_only=False)
    assert set(chk['models']) == {'encoder', 'head'} and set(chk['optimizers']) == {'enc_opt', 'head_opt'}

    import shutil
    shutil.rmtree(eng.work_dir)
    eng = fs_sites(spec=spec)                              # stock learner: only the first network is distributed
    eng.run_nodes(TwoNetTrainer, FSVDataset, max_rounds=2000)
    a, b = eng.site_cache['local0'], eng.site_cache['local1']
    assert eng.trace[-2]['remote'] == 'success' and torch.equal(flat(a, 'encoder'), flat(b, 'encoder'))


def test_custom_data_handle_filters_the_site_listing(fs_sites):
    """"Define custom DataHandle" (reference README, advanced use cases): a user ``COINNDataHandle`` that hides some of a site's
    files from the split (``list_files``) and tweaks the loader arguments is used by the site nodes for everything -
    splitting, training loaders and evaluation."""
    seen = {'listed': 0, 'loaders': 0}

    class EvenSubjectsOnly(COINNDataHandle):
        def list_files(self):
            files = [f for f in super().list_files() if int(f.split('_')[-1].split('.')[0]) % 2 == 0]
            seen['listed'] = len(files)
            return files

        def get_loader(self, handle_key='', **kw):
            seen['loaders'] += 1
            return super().get_loader(handle_key=handle_key, **kw)

    eng = fs_sites(spec={'num_folds': None, 'split_ratio': [0.6, 0.2, 0.2], 'epochs': 1}, sizes=(24, 24))
    eng.run_nodes(FSVTrainer, FSVDataset, datahandle_cls=EvenSubjectsOnly, max_rounds=2000)
    assert eng.trace[-2]['remote'] == 'success' and seen['listed'] == 12 and seen['loaders'] > 0
    split_dir = eng.site_cache['local0']['split_dir']
    with open(os.path.join(split_dir, os.listdir(split_dir)[0])) as fp:
        split = json.load(fp)
    names = split['train'] + split['validation'] + split['test']
    assert len(names) == 12 and all(int(n.split('_')[-1].split('.')[0]) % 2 == 0 for n in names)
