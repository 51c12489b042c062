"""VBMNet - the VBM 3-D CNN classifier (BASELINE.json configs 3 and 4, the headline model).

Architecture (ours to define and document): gray-matter maps 1x121x145x121 ->
5 x [Conv3d(k=3, pad=1, no bias) + BatchNorm3d + ReLU + MaxPool3d(2)] with channels
16-32-64-128-256 -> flatten (256x3x4x3 = 9216) -> Linear(9216, 256) + ReLU -> Linear(256, 64) +
ReLU -> Linear(64, num_class).  ~3.6 M parameters, ~15 GFLOP forward per subject; the first
two blocks are activation-bandwidth bound (68 MB bf16 per subject after conv1), which is what
the fused conv+BN-stat / BN+ReLU+pool kernels target.  The plain-PyTorch twin used by the
reference arm lives in ``baseline/ref_models.py``.
"""
import torch as _torch
from torch import nn as _nn

from .common import ArrayFileDataset, ClassificationTrainer

VBM_INPUT_SHAPE = (1, 121, 145, 121)
VBM_CHANNELS = (16, 32, 64, 128, 256)
VBM_HEAD = (256, 64)


def _pooled(shape, n):
    d, h, w = shape
    for _ in range(n):
        d, h, w = d // 2, h // 2, w // 2
    return d, h, w


class ConvBlock(_nn.Module):
    native_pattern = 'conv_bn_relu_pool'       # ops.nativize: this module IS conv -> bn -> relu -> maxpool(2)

    def __init__(self, cin, cout):
        super().__init__()
        self.conv = _nn.Conv3d(cin, cout, kernel_size=3, padding=1, bias=False)
        self.bn = _nn.BatchNorm3d(cout)
        self.act = _nn.ReLU(inplace=True)
        self.pool = _nn.MaxPool3d(2)

    def forward(self, x):
        return self.pool(self.act(self.bn(self.conv(x))))


class VBMNet(_nn.Module):
    """``native=True`` routes CUDA inputs through the hand-written sm_100a kernels (fused
    conv+BN+ReLU+pool blocks in channels-last bf16, tcgen05 GEMMs for the head) using the very same
    parameters/buffers - ``state_dict`` and optimizer layouts are identical in both modes."""

    def __init__(self, in_ch=1, num_class=2, channels=VBM_CHANNELS, head=VBM_HEAD,
                 input_shape=VBM_INPUT_SHAPE[1:], native=False, conv_backend='auto'):
        super().__init__()
        self.native = bool(native) and in_ch == 1 and channels[0] == 16 and \
            all(c % 8 == 0 and c <= 256 and 256 % (c // 8) == 0 for c in channels)
        self.conv_backend = conv_backend
        chans = [in_ch, *channels]
        self.blocks = _nn.Sequential(*[ConvBlock(a, b) for a, b in zip(chans[:-1], chans[1:])])
        d, h, w = _pooled(input_shape, len(channels))
        feat = channels[-1] * d * h * w
        dims = [feat, *head]
        fc = []
        for a, b in zip(dims[:-1], dims[1:]):
            fc += [_nn.Linear(a, b), _nn.ReLU(inplace=True)]
        self.head = _nn.Sequential(*fc)
        self.classifier = _nn.Linear(dims[-1], num_class)

    @property
    def is_native(self):
        return self.native

    def forward(self, x):
        if self.native and x.is_cuda:
            return self._forward_native(x)
        if x.dim() == 4:
            x = x.unsqueeze(1)
        z = self.blocks(x)
        return self.classifier(self.head(z.flatten(1)))

    def _forward_native(self, x):
        from ..ops.linear import linear as _linear
        from ..ops.vbm import ConvBnReluPoolFn
        h = x[:, 0] if x.dim() == 5 else x                     # [N, D, H, W]; C_in == 1
        if h.dtype not in (_torch.float32, _torch.bfloat16):
            h = h.float()
        # (the fused first block re-lays the volume out as a padded bf16 matrix: fp32 and bf16 volumes go in as they are)
        for blk in self.blocks:
            bn = blk.bn
            h = ConvBnReluPoolFn.apply(h, blk.conv.weight, bn.weight, bn.bias, bn.running_mean, bn.running_var,
                                       bn.eps, bn.momentum if bn.momentum is not None else 0.1, self.training,
                                       self.conv_backend, bn.num_batches_tracked if self.training else None)
        z = h.permute(0, 4, 1, 2, 3).reshape(h.shape[0], -1)    # NCDHW flatten order == the torch path
        for layer in self.head:
            if isinstance(layer, _nn.Linear):
                z = _linear(z, layer.weight, layer.bias, True)
        return _linear(z, self.classifier.weight, self.classifier.bias, False).float()


class VBMDataset(ArrayFileDataset):
    """One ``[121,145,121]`` (or ``[1,121,145,121]``) ``.npy`` volume per subject."""

    def __getitem__(self, ix):
        item = super().__getitem__(ix)
        if item['inputs'].dim() == 3:
            item['inputs'] = item['inputs'].unsqueeze(0)
        return item


class VBMTrainer(ClassificationTrainer):
    def _init_nn_model(self):
        shape = tuple(self.cache.get('input_shape', VBM_INPUT_SHAPE))
        self.nn['vbm_net'] = VBMNet(in_ch=shape[0], num_class=self.cache.get('num_class', 2),
                                    channels=tuple(self.cache.get('channels', VBM_CHANNELS)),
                                    head=tuple(self.cache.get('head', VBM_HEAD)), input_shape=shape[1:],
                                    native=bool(self.cache.get('native_ops', False)),
                                    conv_backend=self.cache.get('conv_backend', 'auto'))
