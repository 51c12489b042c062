from .common import ClassificationTrainer, ArrayFileDataset, InMemorySynthetic, write_synthetic_site  # noqa: F401
from .fsnet import FSNet, FSVDataset, FSVTrainer  # noqa: F401
from .vbmnet import VBMNet, VBMDataset, VBMTrainer, VBM_INPUT_SHAPE  # noqa: F401
