"""Model-selection predicates and a wall-clock recorder (ref utils/utils.py:7-31)."""
import time as _time

from .. import config as _conf

__all__ = ['performance_improved_', 'stop_training_', 'duration']


def performance_improved_(epoch, score, cache):
    """Update ``best_val_score/epoch`` in ``cache`` when ``score`` beats it by ``score_delta``."""
    margin = cache.get('score_delta', _conf.score_delta)
    best = cache['best_val_score']
    direction = cache['metric_direction']
    if direction == 'maximize':
        better = score > best + margin
    elif direction == 'minimize':
        better = score < best - margin
    else:
        better = False
    if better:
        cache['best_val_score'] = score
        cache['best_val_epoch'] = epoch
    return bool(better)


def stop_training_(epoch, cache):
    """Patience rule: stop once ``patience`` epochs passed without improvement."""
    patience = cache.get('patience', cache['epochs'])
    return (epoch - cache['best_val_epoch']) > patience


def duration(cache, begin, key):
    """Append seconds elapsed since ``begin`` to ``cache[key]``; returns the delta (s)."""
    delta = _time.time() - begin
    cache.setdefault(key, [])
    if cache[key] is None:
        cache[key] = []
    cache[key].append(delta)
    return delta
