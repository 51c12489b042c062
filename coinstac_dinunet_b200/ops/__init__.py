"""Hand-written sm_100a kernels and their PyTorch bindings.

``_b200_ops.so`` is built in-tree by ``__graft_entry__.build()`` / ``python -m
coinstac_dinunet_b200.ops.build`` (nvcc, ``arch=compute_100a,code=sm_100a``, ``-lineinfo``) and bound
through ctypes (``native.py``).  On a machine with a GPU the library is REQUIRED: ``native_available``
raises if it is missing, so a silent PyTorch fallback can never masquerade as the native path
(set ``COINN_ALLOW_FALLBACK=1`` to opt out, ``COINN_DISABLE_NATIVE=1`` to force the oracles).  On
CPU-only machines (unit tests, the protocol emulator) the PyTorch oracle implementations are used.
"""
import os as _os

import torch as _torch

_state = {'lib': None, 'err': None}


def _load():
    if _state['lib'] is None and _state['err'] is None:
        try:
            from . import native as _native
            _state['lib'] = _native.lib()
        except Exception as exc:  # noqa: BLE001
            _state['err'] = exc
    return _state['lib']


def extension():
    """The loaded kernel library; raises with the load error if it is missing."""
    lib = _load()
    if lib is None:
        raise RuntimeError(f'coinstac_dinunet_b200.ops._b200_ops.so is not built/loadable: {_state["err"]!r}')
    return lib


def native_available():
    """True when the kernel library is loaded and a CUDA device is present."""
    if _os.environ.get('COINN_DISABLE_NATIVE') == '1' or not _torch.cuda.is_available():
        return False
    if _load() is None:
        if _os.environ.get('COINN_ALLOW_FALLBACK') == '1':
            return False
        raise RuntimeError(f'GPU present but the native kernel library is missing: {_state["err"]!r}')
    return True


#: kernels launched through this package since import (bench.py reports it as ``gpu_launches``)
launch_count = 0


def _count_launch(n=1):
    global launch_count
    launch_count += n


from .api import *  # noqa: E402,F401,F403
from .nativize import nativize  # noqa: E402,F401
