"""FSNet - the FreeSurfer-volumes MLP (BASELINE.json configs 1 and 2).

Architecture (ours to define): 66 aseg volume features ->
[256, 128, 64, 32] x (Linear + BatchNorm1d + ReLU) -> Linear(32, num_class).  ~61 k parameters:
the step is latency-bound, which is what the CUDA-graph + one-shot fused reduce path is for.
The plain-PyTorch twin used by the reference arm lives in ``baseline/ref_models.py``.
"""
from torch import nn as _nn

from .common import ArrayFileDataset, ClassificationTrainer

FS_INPUT_SIZE = 66
FS_HIDDEN = (256, 128, 64, 32)


class FSNet(_nn.Module):
    """``native=True``: every hidden layer (Linear + BatchNorm1d + ReLU) is ONE fused launch forward and ONE backward
    (``ops.linear.linear_bn_relu`` -> ``csrc/linear_small.cu`` at site batch sizes; tcgen05 GEMM + BatchNorm for large
    batches); parameters / buffers / state_dict are unchanged."""

    def __init__(self, in_size=FS_INPUT_SIZE, hidden_sizes=FS_HIDDEN, out_size=2, native=False):
        super().__init__()
        self.native = bool(native)
        dims = [in_size, *hidden_sizes]
        blocks = []
        for a, b in zip(dims[:-1], dims[1:]):
            blocks += [_nn.Linear(a, b), _nn.BatchNorm1d(b), _nn.ReLU(inplace=True)]
        self.features = _nn.Sequential(*blocks)
        self.classifier = _nn.Linear(dims[-1], out_size)

    @property
    def is_native(self):
        return self.native

    def forward(self, x):
        x = x.reshape(x.shape[0], -1)
        if self.native and x.is_cuda:
            from ..ops.linear import linear as _linear, linear_bn_relu as _lbr
            h, mods = x, list(self.features)
            for i in range(0, len(mods), 3):                      # (Linear, BatchNorm1d, ReLU) triples
                h = _lbr(h, mods[i], mods[i + 1], relu=True)
            return _linear(h, self.classifier.weight, self.classifier.bias, False).float()
        return self.classifier(self.features(x))


class FSVDataset(ArrayFileDataset):
    """One 66-float ``.npy`` per subject."""


class FSVTrainer(ClassificationTrainer):
    def _init_nn_model(self):
        self.nn['fs_net'] = FSNet(in_size=self.cache.get('input_size', FS_INPUT_SIZE),
                                  hidden_sizes=tuple(self.cache.get('hidden_sizes', FS_HIDDEN)),
                                  out_size=self.cache.get('num_class', 2),
                                  native=bool(self.cache.get('native_ops', False)))
