"""Global constants (file names of the wire format, metric rounding, device probe).

Parity: coinstac_dinunet/config/__init__.py:5-30.  The file names are part of the
on-disk protocol and therefore identical; everything else is probed
lazily so importing the package never initialises CUDA.
"""
import random as _random
import sys as _sys

from .keys import *  # noqa: F401,F403

# ---- wire / on-disk names -------------------------------------------------
grad_file_ext = '.npy'
grads_file = 'grads' + grad_file_ext
avg_grads_file = 'avg_grads' + grad_file_ext
weights_file = 'weights.tar'

# ---- metrics ---------------------------------------------------------------
metrics_eps = 1e-5
metrics_num_precision = 5

# ---- model selection -------------------------------------------------------
score_delta = 1e-4
score_high = 1.0
score_low = 0.0

max_size = _sys.maxsize

#: Seed drawn once per process; the aggregator adopts it when the user gives none.
current_seed = _random.randint(0, 2 ** 24)


def _probe_cuda():
    try:
        import torch
        return bool(torch.cuda.is_available()), int(torch.cuda.device_count())
    except Exception:  # pragma: no cover - torch is a hard dependency
        return False, 0


CUDA_AVAILABLE, NUM_GPUS = _probe_cuda()

# ---- B200 specifics (used by the NVLink data plane and the bench) ------------
SM_COUNT_B200 = 148
L2_BYTES_B200 = 126 * 1024 * 1024
#: below this bucket size the one-shot (latency) variant of the fused reduce wins
ONE_SHOT_MAX_BYTES = 1 << 20


def boolean_string(s):
    """``'true'`` (any case, surrounding blanks allowed) -> True, anything else -> False."""
    try:
        return str(s).strip().lower() == 'true'
    except Exception:
        return False
