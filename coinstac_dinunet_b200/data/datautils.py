"""Deterministic train/validation/test splitting (k-fold or ratio).

Parity: coinstac_dinunet/data/datautils.py:11-98.  Golden values in SURVEY §8.4 are
unit-tested (tests/test_datautils.py): the shuffle is ``random.seed(len(files))`` so
every site derives the same permutation from the same listing.
"""
import json as _json
import os as _os
import random as _random
import shutil as _shutil

import numpy as _np

_sep = _os.sep


def _seeded_shuffle(files):
    rng = _random.Random(len(files))  # same stream as random.seed(n); random.shuffle(files)
    rng.shuffle(files)
    return files


def _dump(obj, path):
    with open(path, 'w') as fp:
        fp.write(_json.dumps(obj))


def create_ratio_split(files, cache, shuffle_files=True, name='SPLIT'):
    """Split ``files`` by ``cache['split_ratio']`` into train[/validation]/test.

    Two ratios -> (first_key, test); three -> (first_key, validation, test).  Cut points are
    measured *from the end* of the list (so rounding surplus lands in the first key), which is
    what the reference does with its double reversal (datautils.py:27-31).
    """
    ratio = list(cache.get('split_ratio', (0.6, 0.2, 0.2)))
    first_key = cache.get('first_key', 'train')
    out_dir = cache.get('split_dir')
    if shuffle_files:
        _seeded_shuffle(files)

    if len(ratio) == 2:
        keys = [first_key, 'test']
    elif len(ratio) == 3:
        keys = [first_key, 'validation', 'test']
    else:
        keys = [first_key]

    n = len(files)
    # sizes of the trailing parts, last key first
    tail_counts, acc = [], 0.0
    for r in ratio[::-1][:len(keys) - 1]:
        acc += r
        tail_counts.append(int(acc * n))
    bounds = [n - c for c in tail_counts][::-1]  # ascending cut indices
    pieces, lo = [], 0
    for b in bounds + [n]:
        pieces.append(list(files[lo:b]))
        lo = b
    splits = dict(zip(keys, pieces))
    if out_dir:
        _dump(splits, out_dir + _sep + f'{name}.json')
        return None
    return splits


def create_k_fold_splits(files, cache, shuffle_files=True, name='SPLIT'):
    """k folds: fold i tests on chunk i, validates on chunk (i+1) % k, trains on the rest."""
    k = int(cache['num_folds'])
    out_dir = cache.get('split_dir')
    if shuffle_files:
        _seeded_shuffle(files)
    chunks = [c.tolist() for c in _np.array_split(_np.arange(len(files)), k)]
    for i, test_ix in enumerate(chunks):
        val_ix = chunks[(i + 1) % len(chunks)]
        held = set(test_ix) | set(val_ix)
        fold = {
            'train': [files[j] for j in range(len(files)) if j not in held],
            'validation': [files[j] for j in val_ix],
            'test': [files[j] for j in test_ix],
        }
        if not out_dir:
            return fold
        _dump(fold, out_dir + _sep + f'{name}_{i}.json')
    return None


def split_place_holder(files, cache):
    _dump({'train': [], 'validation': [], 'test': []}, cache['split_dir'] + _sep + 'empty_split.json')


def init_k_folds(files, cache, state):
    """Materialise split files under ``<outputDirectory>/<task_id>/splits``.

    Priority (ref datautils.py:73-98): user split directory -> ``split_files`` list ->
    k-fold -> ratio -> empty placeholder.  ``cache['splits']`` maps ``'0','1',..`` to the
    sorted file names.
    """
    user_dir = state['baseDirectory'] + _sep + cache.get('split_dir', 'splits')
    cache['split_dir'] = _os.path.join(state['outputDirectory'], cache['task_id'], 'splits')
    _os.makedirs(cache['split_dir'], exist_ok=True)

    if _os.path.isdir(user_dir) and len(_os.listdir(user_dir)) > 0:
        for f in _os.listdir(user_dir):
            _shutil.copy(user_dir + _sep + f, cache['split_dir'] + _sep + f)
    elif cache.get('split_files'):
        for f in cache['split_files']:
            _shutil.copy(state['baseDirectory'] + _sep + f, cache['split_dir'] + _sep + f)
    elif cache.get('num_folds'):
        create_k_fold_splits(files, cache)
    elif cache.get('split_ratio'):
        create_ratio_split(files, cache)
    else:
        split_place_holder(None, cache)

    names = sorted(_os.listdir(cache['split_dir']))
    cache['splits'] = {str(i): n for i, n in enumerate(names)}
    return {}
