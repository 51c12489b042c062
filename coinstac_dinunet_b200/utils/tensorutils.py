"""Tensor helpers: U-Net skip concat, seeded init, gradient (de)serialisation.

Parity: coinstac_dinunet/utils/tensorutils.py:10-55.  ``save_arrays``/``load_arrays``
are the file wire format: a 1-D numpy object array of per-parameter
ndarrays stored with ``np.save`` (pickle inside ``.npy``).
"""
import numpy as _np
import torch as _torch


def safe_concat(large, small):
    """Center-crop ``large`` to ``small``'s spatial size and concat on channels (4-D / 5-D)."""
    nd = large.dim()
    if nd not in (4, 5):
        raise ValueError(f'safe_concat expects 4-D or 5-D tensors, got {nd}-D')
    slices = [slice(None), slice(None)]
    for ax in range(2, nd):
        extra = large.shape[ax] - small.shape[ax]
        lo = extra // 2
        slices.append(slice(lo, lo + small.shape[ax]))
    return _torch.cat([large[tuple(slices)], small], dim=1)


_KAIMING_TYPES = (_torch.nn.Conv2d, _torch.nn.Linear, _torch.nn.Conv3d, _torch.nn.Conv1d)
_UNIT_NORM_TYPES = (_torch.nn.BatchNorm1d, _torch.nn.BatchNorm2d, _torch.nn.BatchNorm3d)


def initialize_weights(*models, extended=True):
    """Kaiming-normal weights / zero bias; BatchNorm -> (1, 0).

    The reference touches only Conv2d / Linear / BatchNorm2d
    (tensorutils.py:28-37, quirk 8.5-10).  ``extended=True`` (our default) also covers
    Conv1d/Conv3d and BatchNorm1d/3d so the VBM 3-D CNN is seeded identically on every
    site; ``extended=False`` reproduces the reference exactly (used by the golden tests).
    """
    conv_t = _KAIMING_TYPES if extended else (_torch.nn.Conv2d, _torch.nn.Linear)
    norm_t = _UNIT_NORM_TYPES if extended else (_torch.nn.BatchNorm2d,)
    for model in models:
        for m in model.modules():
            if isinstance(m, conv_t):
                _torch.nn.init.kaiming_normal_(m.weight)
                if m.bias is not None:
                    with _torch.no_grad():
                        m.bias.zero_()
            elif isinstance(m, norm_t) and m.weight is not None:
                with _torch.no_grad():
                    m.weight.fill_(1)
                    m.bias.zero_()


def caste_ndarray(a, dtype='float32'):
    return a.astype(dtype)


def extract_grads(model, dtype='float32'):
    """Host copies of every parameter gradient in ``model.parameters()`` order."""
    out = []
    for p in model.parameters():
        g = p.grad
        if g is None:  # quirk 8.5-16: the reference crashes here; we ship zeros instead
            g = _torch.zeros_like(p)
        out.append(caste_ndarray(g.detach().float().cpu().numpy(), dtype))
    return out


def as_object_array(arrays):
    """1-D object array holding ``arrays`` regardless of their shapes (SURVEY §8.1 caveat)."""
    if isinstance(arrays, _np.ndarray) and arrays.dtype == object and arrays.ndim == 1:
        return arrays
    box = _np.empty(len(arrays), dtype=object)
    for i, a in enumerate(arrays):
        box[i] = a
    return box


def save_arrays(file_path, arrays):
    if not (isinstance(arrays, _np.ndarray) and arrays.dtype != object):
        arrays = as_object_array(list(arrays))
    with open(file_path, 'wb') as fp:  # explicit handle: np.save would append '.npy' to odd names
        _np.save(fp, arrays, allow_pickle=True)


def load_arrays(file_path):
    return _np.load(file_path, allow_pickle=True)
