"""Protocol vocabulary: phases, modes, cache/JSON key names and engine selectors.

The string values are the wire format of the COINSTAC JSON control plane and must
stay byte-compatible with the reference (coinstac_dinunet/config/keys.py:4-49).
All enums derive from ``str`` so they JSON-serialise and compare equal to the raw
strings a real COINSTAC engine ships around.
"""
from enum import Enum

__all__ = ['Phase', 'Mode', 'Key', 'AGG_Engine', 'GatherMode', 'Transport']


class _Wire(str, Enum):
    """str-valued enum whose ``str()``/``format()`` is the raw wire value."""

    def __str__(self):
        return str(self.value)

    def __format__(self, spec):
        return format(str(self.value), spec)


class Phase(_Wire):
    """Life-cycle of one fold (ref keys.py:4-10)."""
    INIT_RUNS = 'init_runs'                 # sites make splits, share args
    NEXT_RUN = 'next_run'                   # remote hands out fold + seed
    PRE_COMPUTATION = 'pre_computation'     # optional local pre-training / weight broadcast
    COMPUTATION = 'computation'             # distributed train / validation / test
    NEXT_RUN_WAITING = 'next_run_waiting'   # site finished its test pass
    SUCCESS = 'success'                     # all folds done, results zip shipped


class Mode(_Wire):
    """Per-site activity inside Phase.COMPUTATION (ref keys.py:13-19)."""
    PRE_TRAIN = 'pre_train'
    TRAIN = 'train'
    VALIDATION = 'validation'
    TEST = 'test'
    VALIDATION_WAITING = 'validation_waiting'
    TRAIN_WAITING = 'train_waiting'


class Key(_Wire):
    """Names of log / score slots in ``cache`` and in the JSON payloads (ref keys.py:22-38)."""
    ARGS_CACHED = '_args_cached_'

    TRAIN_LOG = 'train_log'
    TRAIN_METRICS = 'train_metrics'
    TRAIN_SERIALIZABLE = 'serializable_train_scores'

    VALIDATION_LOG = 'validation_log'
    VALIDATION_METRICS = 'validation_metrics'
    VALIDATION_SERIALIZABLE = 'serializable_validation_scores'

    TEST_LOG = 'test_log'
    TEST_METRICS = 'test_metrics'
    TEST_SERIALIZABLE = 'serializable_test_scores'

    GLOBAL_TEST_LOG = 'global_test_log'
    GLOBAL_TEST_METRICS = 'global_test_metrics'
    GLOBAL_TEST_SERIALIZABLE = 'serializable_global_test_scores'


class AGG_Engine(_Wire):
    """Aggregation engines selectable through ``cache['agg_engine']`` (ref keys.py:41-44)."""
    dSGD = 'dSGD'
    powerSGD = 'powerSGD'
    rankDAD = 'rankDAD'


class GatherMode(_Wire):
    """How the aggregator collects per-site lists (ref keys.py:47-49)."""
    APPEND = 'gather'
    EXTEND = 'extend'


class Transport(_Wire):
    """Data-plane used for the per-step exchange (new in this framework).

    * ``FILE``   - reference-compatible ``*.npy`` + JSON round trip through an engine.
    * ``NVLINK`` - in-kernel peer loads/stores over NVSwitch, fused with the optimizer.
    * ``NCCL``   - torch.distributed all-reduce (the baseline, kept for A/B measurements).
    """
    FILE = 'file'
    NVLINK = 'nvlink'
    NCCL = 'nccl'
