from .metrics import COINNMetrics, COINNAverages, Prf1a, ConfusionMatrix, AUCROCMetrics  # noqa: F401
from .loss import dice_loss_binary  # noqa: F401
from . import metrics, loss  # noqa: F401
