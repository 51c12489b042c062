"""Small host-side utilities: FrozenDict, CSV/JSON dumps of the run cache.

Parity: coinstac_dinunet/utils/__init__.py:8-80.
"""
import copy as _copy
import json as _json
import os as _os

from .logger import lazy_debug  # noqa: F401  (re-exported like the reference)
from .utils import *  # noqa: F401,F403


class FrozenDict(dict):
    """Write-once mapping: new keys may be added, existing keys can never be rebound.

    Used to protect ``input`` / ``state`` / shared args from accidental mutation
    (ref utils/__init__.py:8-26).
    """

    def __init__(self, _dict=None):
        super().__init__(_dict if _dict is not None else {})

    def prompt(self, key, value):
        raise ValueError(
            f"*** Attempt to modify frozen dict [{key} : {self[key]}] with [{key} : {value}] ***")

    def __setitem__(self, key, value):
        if key in self:
            self.prompt(key, value)
        super().__setitem__(key, value)

    def update(self, *args, **kw):
        for src in args:
            for k, v in dict(src).items():
                self[k] = v
        for k, v in kw.items():
            self[k] = v

    def setdefault(self, key, default=None):
        if key not in self:
            super().__setitem__(key, default)
        return self[key]


def save_scores(cache, log_dir, file_keys=()):
    """One ``<key>.csv`` per key; first line is ``log_header`` (ref utils/__init__.py:29-40)."""
    header = cache.get('log_header', '')
    if isinstance(header, (list, tuple)):
        header = ','.join(header)
    for fk in file_keys:
        rows = cache[fk]
        with open(_os.path.join(log_dir, f'{fk}.csv'), 'w') as fp:
            fp.write(f"{header}\n")
            for row in rows:
                if isinstance(row, (list, tuple)):
                    fp.write(','.join(str(v) for v in row) + '\n')
                else:
                    fp.write(f"{row}\n")


def jsonable(obj):
    try:
        _json.dumps(obj)
    except Exception:
        return False
    return True


def clean_recursive(obj):
    """In-place: replace every non-JSON leaf of a nested dict by its ``str()``."""
    if not isinstance(obj, dict):
        return
    for k in list(obj.keys()):
        v = obj[k]
        if isinstance(v, dict):
            clean_recursive(v)
        elif isinstance(v, list):
            for item in v:
                clean_recursive(item)
        elif not jsonable(v):
            obj[k] = f'{v}'


def save_cache(cache, log_dir):
    """Dump a JSON-safe snapshot of ``cache`` to ``<log_dir>/logs.json``.

    Live objects (modules, optimizers, iterators, device arenas) are stringified;
    entries that cannot even be deep-copied become ``''`` (ref utils/__init__.py:67-76).
    """
    snapshot = {}
    for k in cache.keys():
        if str(k).startswith('_') and str(k) != '_args_cached_':
            continue          # private runtime objects (device arenas, CUDA graphs, streams): never worth a deep copy
        try:
            snapshot[str(k)] = _copy.deepcopy(cache[k])
        except Exception:
            snapshot[str(k)] = ''
    clean_recursive(snapshot)
    with open(_os.path.join(log_dir, 'logs.json'), 'w') as fp:
        _json.dump(snapshot, fp)
