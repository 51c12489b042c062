"""Aggregator entry point of the user-defined-model computation."""
from multiprocessing.pool import ThreadPool

from coinstac_dinunet_b200 import COINNRemote

try:
    from .local import MyTrainer
except ImportError:                     # executed as a script
    from local import MyTrainer

_cache, _pool = {}, None


def compute(args):
    global _pool
    _pool = _pool or ThreadPool(2)
    cache = args.get('cache') if args.get('cache') is not None else _cache
    return COINNRemote(cache=cache, input=args['input'], state=args['state'])(_pool, MyTrainer)


if __name__ == '__main__':
    try:
        import coinstac
        coinstac.start(None, compute)
    except ImportError:
        raise SystemExit('run under COINSTAC')
