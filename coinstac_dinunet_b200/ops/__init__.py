"""Hand-written sm_100a kernels and their PyTorch bindings.

The extension ``_b200_ops`` is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a,
``-lineinfo``).  On a machine with a GPU the extension is REQUIRED: importing the ops without it
raises, so a silent PyTorch fallback can never masquerade as the native path.  On CPU-only
machines (unit tests, the protocol emulator) the pure-PyTorch reference implementations in
``ops.reference`` are used.
"""
import os as _os

import torch as _torch

_ext = None
_ext_err = None


def _load():
    global _ext, _ext_err
    if _ext is not None or _ext_err is not None:
        return _ext
    try:
        from . import _b200_ops as ext  # built in-tree: coinstac_dinunet_b200/ops/_b200_ops*.so
        _ext = ext
    except Exception as exc:  # noqa: BLE001
        _ext_err = exc
    return _ext


def extension():
    """The loaded extension module; raises with the import error if it is missing."""
    ext = _load()
    if ext is None:
        raise RuntimeError(
            'coinstac_dinunet_b200.ops._b200_ops is not built/loaded '
            f'(run `python -c "import __graft_entry__ as g; g.build()"`): {_ext_err!r}')
    return ext


def native_available():
    """True when the extension is loaded and a CUDA device is present."""
    if _os.environ.get('COINN_DISABLE_NATIVE') == '1':
        return False
    if not _torch.cuda.is_available():
        return False
    if _load() is None:
        if _os.environ.get('COINN_ALLOW_FALLBACK') == '1':
            return False
        raise RuntimeError(f'GPU present but native extension missing: {_ext_err!r}')
    return True


from .api import *  # noqa: E402,F401,F403
