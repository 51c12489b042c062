"""Reference arm: drive the UNMODIFIED trendscenter/coinstac-dinunet (installed in
``baseline/_ref``) through its own public API - COINNLocal / COINNRemote / COINNTrainer /
COINNDataset / COINNDataHandle, stock dSGD learner + reducer, ``grads.npy`` / ``avg_grads.npy``
files - with a minimal stand-in for the external COINSTAC engine: one process per
site (rank r == site ``local<r>``), rank 0 also hosts the aggregator; JSON via
``torch.distributed`` object collectives, files copied on the local filesystem.

Nothing from coinstac_dinunet_b200 is imported here.  Two environment shims are needed because
the reference does not import on this image - they patch the *environment*, not
the reference: a no-op ``matplotlib`` stub (package not installed) and the ``np.float``/``np.int``
aliases NumPy 2 removed.
"""
import os
import shutil
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, '_ref')


def _install_shims():
    import numpy as np
    for name, typ in (('float', float), ('int', int), ('bool', bool)):
        if not hasattr(np, name):
            setattr(np, name, typ)
    try:
        import matplotlib  # noqa: F401
    except Exception:
        mpl = types.ModuleType('matplotlib')
        plt = types.ModuleType('matplotlib.pyplot')

        class _Noop:
            def __getattr__(self, _):
                return lambda *a, **k: _Noop()

            def __call__(self, *a, **k):
                return _Noop()

        plt.rcParams = {}
        for fn in ('switch_backend', 'clf', 'xlabel', 'savefig', 'close', 'figure', 'plot'):
            setattr(plt, fn, lambda *a, **k: None)
        mpl.pyplot = plt
        mpl.use = lambda *a, **k: None
        sys.modules['matplotlib'] = mpl
        sys.modules['matplotlib.pyplot'] = plt
        # pandas' DataFrame.plot would need matplotlib: make plot_progress a no-op downstream
        os.environ['COINN_REF_NO_PLOTS'] = '1'


def import_reference():
    """Returns the reference package or raises ImportError with the reason."""
    if not os.path.isdir(os.path.join(REF_DIR, 'coinstac_dinunet')):
        raise ImportError(f'reference not installed under {REF_DIR}')
    _install_shims()
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import coinstac_dinunet  # noqa: F401
    import coinstac_dinunet.vision.plotter as plotter
    if os.environ.get('COINN_REF_NO_PLOTS') == '1':
        # matplotlib is absent on this image; plotting is not part of the measured step
        plotter.plot_progress = lambda *a, **k: None
        import coinstac_dinunet.nn.basetrainer as bt
        import coinstac_dinunet.distrib.nodes.remote as rm
        bt._plot.plot_progress = plotter.plot_progress
        rm._plot.plot_progress = plotter.plot_progress
    return coinstac_dinunet


# ----------------------------------------------------------------------------- user code
def build_user_classes(model_name, input_shape, num_class=2):
    """What a user of the reference writes: a dataset, a trainer (README.md:41-84)."""
    import torch
    import torch.nn.functional as F
    from coinstac_dinunet import COINNDataset, COINNTrainer
    sys.path.insert(0, HERE)
    from ref_models import RefFSNet, RefVBMNet

    class SyntheticDataset(COINNDataset):
        """Synthetic subjects of the named shape; the file name only seeds the generator."""
        _pool = {}

        def __getitem__(self, ix):
            file = self.indices[ix][0]
            key = hash(file) % 64            # 64 distinct volumes are plenty for a throughput run
            if key not in self._pool:
                g = torch.Generator().manual_seed(key)
                self._pool[key] = (torch.randn(*input_shape, generator=g), int(key % num_class))
            x, y = self._pool[key]
            return {'inputs': x, 'labels': torch.tensor(y)}

    class RefTrainer(COINNTrainer):
        def _init_nn_model(self):
            if model_name == 'vbm':
                self.nn['model'] = RefVBMNet(in_ch=input_shape[0], num_class=num_class, input_shape=input_shape[1:])
            else:
                self.nn['model'] = RefFSNet(in_size=input_shape[0], out_size=num_class)

        def iteration(self, batch):
            inputs = batch['inputs'].to(self.device['gpu']).float()
            labels = batch['labels'].to(self.device['gpu']).long()
            out = F.log_softmax(self.nn['model'](inputs), 1)
            loss = F.nll_loss(out, labels)
            _, predicted = torch.max(out, 1)
            score = self.new_metrics()
            score.add(predicted, labels)
            val = self.new_averages()
            val.add(loss.item(), len(inputs))
            return {'out': out, 'loss': loss, 'averages': val, 'metrics': score, 'prediction': predicted}

    return SyntheticDataset, RefTrainer


# ----------------------------------------------------------------------------- mini engine
def _state(work, node):
    st = {'clientId': node}
    for key, top in (('baseDirectory', 'input'), ('outputDirectory', 'output'), ('transferDirectory', 'transfer')):
        st[key] = os.path.join(work, top, node, 'simulatorRun')
        os.makedirs(st[key], exist_ok=True)
    return st


def _copy_files(src, dst):
    os.makedirs(dst, exist_ok=True)
    for f in os.listdir(src):
        p = os.path.join(src, f)
        if os.path.isfile(p):
            shutil.copy(p, os.path.join(dst, f))


def _plain(o):
    if isinstance(o, dict):
        return {str(getattr(k, 'value', k)): _plain(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return [_plain(v) for v in o]
    return getattr(o, 'value', o) if isinstance(o, str) else o


class RefEngine:
    """rank r runs site local<r>; rank 0 also runs the remote.  One ``round()`` == one dSGD step
    in the computation/train phase."""

    def __init__(self, work, model_name, input_shape, batch_size, n_files, use_gpu, num_class=2):
        import torch.distributed as dist
        from multiprocessing.pool import ThreadPool
        self.dist = dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.site = f'local{self.rank}'
        self.state = _state(work, self.site)
        self.sites = [f'local{i}' for i in range(self.world)]
        self.cache, self.remote_cache = {}, {}
        self.remote_state = _state(work, 'remote') if self.rank == 0 else None
        self.all_states = [_state(work, s) for s in self.sites] if self.rank == 0 else None
        self.pool = ThreadPool(2)
        import atexit
        atexit.register(self.pool.terminate)
        self.dataset_cls, self.trainer_cls = build_user_classes(model_name, input_shape, num_class)
        data_dir = os.path.join(self.state['baseDirectory'], 'data')
        os.makedirs(data_dir, exist_ok=True)
        for i in range(n_files):
            open(os.path.join(data_dir, f's{self.rank:02d}_{i:06d}'), 'w').close()
        local_rank = int(os.environ.get('LOCAL_RANK', self.rank))
        self.input = {'task_id': 'bench', 'mode': 'train', 'data_dir': 'data', 'num_class': num_class,
                      'batch_size': batch_size, 'split_ratio': [0.98, 0.01, 0.01], 'epochs': 10 ** 6,
                      'gpus': [local_rank] if use_gpu else None, 'agg_engine': 'dSGD', 'seed': 11,
                      'monitor_metric': 'f1', 'precision_bits': 32}
        self.num_class = num_class
        self.rounds = 0

    def round(self):
        from coinstac_dinunet import COINNLocal, COINNRemote
        from coinstac_dinunet.data import COINNDataHandle
        dist = self.dist
        node = COINNLocal(cache=self.cache, input=self.input, state=self.state)
        out = node(self.pool, self.trainer_cls, self.dataset_cls, COINNDataHandle)['output']
        gathered = [None] * self.world if self.rank == 0 else None
        dist.gather_object(_plain(out), gathered, dst=0)
        payload = [None]
        if self.rank == 0:
            for st in self.all_states:
                _copy_files(st['transferDirectory'], os.path.join(self.remote_state['baseDirectory'], st['clientId']))
            rnode = COINNRemote(cache=self.remote_cache, input=dict(zip(self.sites, gathered)),
                                state=self.remote_state, num_class=self.num_class)
            res = rnode(self.pool, self.trainer_cls)
            for st in self.all_states:
                _copy_files(self.remote_state['transferDirectory'], st['baseDirectory'])
            payload = [(_plain(res['output']), bool(res.get('success')))]
        dist.broadcast_object_list(payload, src=0)
        self.input = payload[0][0]
        self.rounds += 1
        if payload[0][1]:
            raise RuntimeError('reference run finished (success) - dataset too small for the requested steps')
        return out, self.input

    def advance_to_training(self, max_rounds=10):
        """init_runs -> next_run -> first computation round (gradients flowing)."""
        for _ in range(max_rounds):
            out, rin = self.round()
            if rin.get('update'):
                return
        raise RuntimeError('reference never reached the training phase')
