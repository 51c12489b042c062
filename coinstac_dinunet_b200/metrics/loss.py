"""Segmentation losses (ref metrics/loss.py:1-22)."""
import torch as _torch


def dice_loss_binary(outputs=None, target=None, beta=1, weights=None):
    r"""Weighted F-beta / Dice loss with unit smoothing.

    ``1 - ((1+b^2)·Σ w·o·t + 1) / (b^2·Σ w·o + Σ w·t + 1)``.  ``weights`` with a zero
    minimum are shifted by the smoothing constant so no voxel is ignored entirely.
    All three sums come from one pass over the flattened tensors (``torch.stack`` +
    a single reduction) instead of three separate reductions.
    """
    smooth = 1.0
    o = outputs.contiguous().float().reshape(-1)
    t = target.contiguous().float().reshape(-1)
    if weights is not None:
        w = weights.contiguous().float().reshape(-1)
        w = _torch.where(w.min() == 0, w + smooth, w)
        terms = _torch.stack([o * t * w, w * o, w * t]).sum(1)
    else:
        terms = _torch.stack([o * t, o, t]).sum(1)
    inter, so, st = terms[0], terms[1], terms[2]
    b2 = beta ** 2
    return 1.0 - ((1 + b2) * inter + smooth) / (b2 * so + st + smooth)
