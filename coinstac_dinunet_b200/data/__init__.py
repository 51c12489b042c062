from .data import (COINNDataset, COINNDataHandle, COINNPaddedDataSampler, DevicePrefetcher,  # noqa: F401
                   safe_collate)
from . import datautils  # noqa: F401
