"""coinstac_dinunet_b200 - a Blackwell-native federated / distributed-SGD training engine
with the public API of trendscenter/coinstac-dinunet.

    from coinstac_dinunet_b200 import COINNDataset, COINNDataHandle, COINNTrainer, \
        COINNLocal, COINNRemote, COINNLearner, COINNReducer

Sub-packages: ``config`` ``utils`` ``metrics`` ``data`` ``nn`` ``distrib`` ``vision`` (reference
parity) and ``engine`` ``parallel`` ``ops`` ``models`` (the B200 runtime: in-process / per-GPU
round engines, symmetric-memory arenas + fused reduce/optimizer kernels, hand-written sm_100a
ops, and the FreeSurfer-MLP / VBM-3D-CNN model zoo).
"""
try:
    import torch as _torch  # noqa: F401
except Exception as _exc:  # pragma: no cover
    raise ImportError(
        'coinstac_dinunet_b200 needs PyTorch (a CUDA 12.8+ build for the sm_100a kernels).') from _exc

__version__ = '0.1.0'

from .data import COINNDataset, COINNDataHandle  # noqa: E402,F401
from .distrib import COINNLearner, COINNReducer  # noqa: E402,F401
from .distrib import COINNLocal, COINNRemote  # noqa: E402,F401
from .trainer import COINNTrainer  # noqa: E402,F401
from .site_runner import SiteRunner  # noqa: E402,F401
