"""CPU image / patch utilities for 2-D segmentation pipelines.

Function-for-function parity with coinstac_dinunet/vision/imageutils.py:21-348 (same names,
arguments and return conventions); none of this is on a training hot path.
OpenCV, scikit-image and SciPy are imported lazily and have NumPy fallbacks, so the module
imports on a bare image.
"""
import copy as _copy
import math as _math
import os as _os

import numpy as _np


def _pil():
    from PIL import Image as _IMG
    return _IMG


def _read(path, dtype):
    return _np.array(_pil().open(path), dtype=dtype)


class Image:
    """An image with optional mask / ground truth and a free-form ``extras`` dict."""

    def __init__(self, dtype=_np.uint8):
        self.dir = None
        self.file = None
        self.array = None
        self.mask = None
        self.ground_truth = None
        self.extras = {}
        self.dtype = dtype

    @property
    def path(self):
        return _os.path.join(self.dir, self.file)

    def load(self, dir, file):
        self.dir, self.file = dir, file
        try:
            self.array = _read(self.path, self.dtype)
        except Exception as exc:
            print(f'### Error Loading file: {self.file}: {exc}')

    def load_mask(self, mask_dir=None, fget_mask=lambda x: x):
        try:
            self.mask = _read(_os.path.join(mask_dir, fget_mask(self.file)), self.dtype)
        except Exception as exc:
            print(f'### Fail to load mask: {exc}')

    def load_ground_truth(self, gt_dir=None, fget_ground_truth=lambda x: x):
        try:
            self.ground_truth = _read(_os.path.join(gt_dir, fget_ground_truth(self.file)), self.dtype)
        except Exception as exc:
            print(f'### Fail to load ground truth: {exc}')

    def get_array(self, dir='', getter=lambda x: x, file=None):
        return _read(_os.path.join(dir, getter(file if file else self.file)), self.dtype)

    def apply_mask(self):
        if self.mask is not None:
            self.array[self.mask == 0] = 0

    def apply_clahe(self, clip_limit=2.0, tile_shape=(8, 8)):
        import cv2
        clahe = cv2.createCLAHE(clipLimit=clip_limit, tileGridSize=tile_shape)
        if self.array.ndim == 2:
            self.array = clahe.apply(self.array)
        elif self.array.ndim == 3 and self.array.shape[2] >= 3:
            for ch in range(3):
                self.array[:, :, ch] = clahe.apply(self.array[:, :, ch])
        else:
            print('### More than three channels')

    def __copy__(self):
        twin = Image(dtype=_copy.deepcopy(self.dtype))
        twin.dir = self.dir
        twin.file = _copy.copy(self.file)
        twin.array = _copy.copy(self.array)
        twin.mask = _copy.copy(self.mask)
        twin.ground_truth = _copy.copy(self.ground_truth)
        twin.extras = _copy.deepcopy(self.extras)
        return twin


def _binary_codes(arr_2d, truth):
    """``pred + 2*true`` with 255 -> 1: 0 TN, 1 FP, 2 FN, 3 TP."""
    p = _np.where(arr_2d == 255, 1, arr_2d).astype(_np.int64)
    t = _np.where(truth == 255, 1, truth).astype(_np.int64)
    return p + 2 * t


def get_rgb_scores(arr_2d=None, truth=None):
    """Colour-coded agreement map: TP white, FP green, FN red, TN black."""
    codes = _binary_codes(arr_2d, truth)
    palette = _np.array([[0, 0, 0], [0, 255, 0], [255, 0, 0], [255, 255, 255]], dtype=_np.uint8)
    return palette[_np.clip(codes, 0, 3)]


def get_praf1(arr_2d=None, truth=None):
    """Precision / Recall / Accuracy / F1 (5 decimals) between two binary arrays."""
    counts = _np.bincount(_np.clip(_binary_codes(arr_2d, truth), 0, 3).reshape(-1), minlength=4)
    tn, fp, fn, tp = (int(c) for c in counts[:4])

    def ratio(a, b):
        return a / b if b else 0

    p, r = ratio(tp, tp + fp), ratio(tp, tp + fn)
    return {'Precision': round(p, 5), 'Recall': round(r, 5),
            'Accuracy': round(ratio(tp + tn, tp + fp + fn + tn), 5),
            'F1': round(ratio(2 * p * r, p + r), 5)}


def rescale2d(arr):
    lo, hi = _np.min(arr), _np.max(arr)
    return (arr - lo) / (hi - lo)


def rescale3d(arrays):
    return [rescale2d(a) for a in arrays]


def get_signed_diff_int8(image_arr1=None, image_arr2=None):
    diff = _np.array(image_arr1 - image_arr2, dtype=_np.int8)
    shifted = _np.array(diff - _np.min(diff), _np.uint8)
    return _np.array(rescale2d(shifted) * 255, _np.uint8)


def whiten_image2d(img_arr2d=None):
    z = (img_arr2d - img_arr2d.mean()) / img_arr2d.std()
    return _np.array(rescale2d(z) * 255, dtype=_np.uint8)


def _axis_windows(length, size, stride):
    """Start/stop pairs along one axis; the last window is shifted back to stay inside."""
    spans = []
    for start in range(0, length, stride):
        stop = start + size
        if stop > length:
            spans.append((length - size, length))
            break
        spans.append((start, stop))
    return spans


def get_chunk_indexes(img_shape=(0, 0), chunk_shape=(0, 0), offset_row_col=None):
    """Yield ``[row_from, row_to, col_from, col_to]`` for every patch, row-major."""
    rows = _axis_windows(img_shape[0], chunk_shape[0], offset_row_col[0])
    cols = _axis_windows(img_shape[1], chunk_shape[1], offset_row_col[1])
    for r0, r1 in rows:
        for c0, c1 in cols:
            yield [int(r0), int(r1), int(c0), int(c1)]


def get_chunk_indices_by_index(img_shape=(0, 0), chunk_shape=(0, 0), indices=None):
    """Patches of ``chunk_shape`` centred on ``indices``, clamped to the image."""
    h, w = chunk_shape
    H, W = img_shape

    def clamp(center, size, limit):
        lo, hi = center - size // 2, center + size // 2
        if lo < 0:
            lo, hi = 0, size
        if hi > limit:
            lo, hi = limit - size, limit
        return int(lo), int(hi)

    out = []
    for ci, cj in indices:
        p, q = clamp(ci, h, H)
        r, s = clamp(cj, w, W)
        out.append([p, q, r, s])
    return out


def merge_patches(patches=None, image_size=(0, 0), patch_size=(0, 0), offset_row_col=None):
    """Stitch patches back; overlaps are averaged over the patches that are non-zero there."""
    total = _np.zeros(tuple(image_size[:2]), dtype=_np.float64)
    hits = _np.zeros_like(total)
    for i, (r0, r1, c0, c1) in enumerate(get_chunk_indexes(image_size, patch_size, offset_row_col)):
        patch = _np.array(patches[i, :, :]).squeeze()
        total[r0:r1, c0:c1] += patch
        hits[r0:r1, c0:c1] += (patch > 0)
    hits[hits == 0] = 1
    return _np.array(total / hits, dtype=_np.uint8)


def expand_and_mirror_patch(full_img_shape=None, orig_patch_indices=None, expand_by=None):
    """Grow a patch by ``expand_by`` (split evenly); returns the clipped corners and the
    per-side padding needed to mirror what fell outside the image."""
    half_r, half_c = int(expand_by[0] / 2), int(expand_by[1] / 2)
    p, q, r, s = orig_patch_indices
    a, b, c, d = p - half_r, q + half_r, r - half_c, s + half_c
    pad_a = pad_b = pad_c = pad_d = 0
    if a < 0:
        pad_a, a = half_r - p, 0
    if b > full_img_shape[0]:
        pad_b, b = b - full_img_shape[0], full_img_shape[0]
    if c < 0:
        pad_c, c = half_c - r, 0
    if d > full_img_shape[1]:
        pad_d, d = d - full_img_shape[1], full_img_shape[1]
    return a, b, c, d, [(pad_a, pad_b), (pad_c, pad_d)]


def _label(binary, structure=None):
    from scipy import ndimage
    return ndimage.label(binary, structure)


def largest_cc(binary_arr=None):
    """Mask of the largest connected component (``None`` when there is none)."""
    try:
        from skimage.measure import label as sk_label
        labels = sk_label(binary_arr)
    except Exception:
        labels, _ = _label(binary_arr, _np.ones((3,) * _np.ndim(binary_arr), dtype=int))
    if labels.max() != 0:
        return labels == _np.argmax(_np.bincount(labels.flat)[1:]) + 1


def map_img_to_img2d(map_to, img):
    """Paint pixels where ``img == 255`` in red on (a grey->RGB copy of) ``map_to``."""
    base = map_to.copy()
    rgb = _np.stack([base] * 3, axis=-1).astype(_np.uint8) if base.ndim == 2 else base.copy()
    hit = img == 255
    rgb[hit, 0], rgb[hit, 1], rgb[hit, 2] = 255, 0, 0
    return rgb


def remove_connected_comp(segmented_img, connected_comp_diam_limit=20):
    """Erase 8-connected components whose first-to-last pixel distance is below the limit."""
    out = segmented_img.copy()
    labeled, n = _label(out, _np.ones((3, 3), dtype=int))
    for lab in range(n):
        pts = _np.argwhere(labeled == lab)
        if len(pts) == 0:
            continue
        (x1, y1), (x2, y2) = pts[0], pts[-1]
        if _math.hypot(x2 - x1, y2 - y1) < connected_comp_diam_limit:
            out[pts[:, 0], pts[:, 1]] = 0
    return out


def get_pix_neigh(i, j, eight=False):
    """4- (N, E, S, W) or 8-neighbourhood (row-major) of pixel ``(i, j)``."""
    if eight:
        return [(i + di, j + dj) for di in (-1, 0, 1) for dj in (-1, 0, 1) if (di, dj) != (0, 0)]
    return [(i - 1, j), (i, j + 1), (i + 1, j), (i, j - 1)]
