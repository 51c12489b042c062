"""Site entry point of a computation with a USER-DEFINED model: the reference's core use case
(`self.nn[...] = <your nn.Module>`, coinstac_dinunet/nn/basetrainer.py:30-34).  With `"native_ops": true` in the
inputspec the trainer calls `ops.nativize` on the module: the Conv3d-BN-ReLU-MaxPool and Linear-BN1d-ReLU runs of the
`nn.Sequential`s below execute on the hand-written sm_100a kernels (channel counts the kernels are not instantiated for are
zero-padded), parameters / `state_dict` stay exactly what PyTorch would have."""
from multiprocessing.pool import ThreadPool

from torch import nn

from coinstac_dinunet_b200 import COINNLocal
from coinstac_dinunet_b200.models import ClassificationTrainer, VBMDataset


class MyNet(nn.Module):
    def __init__(self, in_shape=(64, 64, 64), num_class=2):
        super().__init__()
        self.features = nn.Sequential(
            nn.Conv3d(1, 8, 3, padding=1), nn.BatchNorm3d(8), nn.ReLU(), nn.MaxPool3d(2),
            nn.Conv3d(8, 24, 3, padding=1, bias=False), nn.BatchNorm3d(24), nn.ReLU(), nn.MaxPool3d(2),
            nn.Conv3d(24, 48, 3, padding=1, bias=False), nn.BatchNorm3d(48), nn.ReLU(), nn.MaxPool3d(2))
        feat = 48 * (in_shape[0] // 8) * (in_shape[1] // 8) * (in_shape[2] // 8)
        self.head = nn.Sequential(nn.Flatten(), nn.Linear(feat, 64), nn.BatchNorm1d(64), nn.ReLU(), nn.Linear(64, num_class))

    def forward(self, x):
        return self.head(self.features(x))


class MyTrainer(ClassificationTrainer):
    def _init_nn_model(self):
        self.nn['my_net'] = MyNet(tuple(self.cache.get('input_shape', (1, 64, 64, 64)))[1:], self.cache.get('num_class', 2))


_cache, _pool = {}, None


def compute(args):
    global _pool
    _pool = _pool or ThreadPool(2)
    cache = args.get('cache') if args.get('cache') is not None else _cache
    node = COINNLocal(cache=cache, input=args['input'], state=args['state'], task_id='custom', epochs=11, batch_size=8,
                      learning_rate=1e-3, monitor_metric='f1')
    return node(_pool, MyTrainer, VBMDataset)


if __name__ == '__main__':
    try:
        import coinstac
        coinstac.start(compute, None)
    except ImportError:
        raise SystemExit('run under COINSTAC, or drive it with coinstac_dinunet_b200.engine.{InProcessEngine,DistEngine}')
