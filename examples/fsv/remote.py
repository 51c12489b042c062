"""Aggregator entry point of the FreeSurfer-volumes computation (COINSTAC remote node)."""
from multiprocessing.pool import ThreadPool

from coinstac_dinunet_b200 import COINNRemote
from coinstac_dinunet_b200.models import FSVTrainer

_cache, _pool = {}, None


def compute(args):
    global _pool
    _pool = _pool or ThreadPool(2)
    cache = args.get('cache') if args.get('cache') is not None else _cache
    return COINNRemote(cache=cache, input=args['input'], state=args['state'])(_pool, FSVTrainer)


if __name__ == '__main__':
    try:
        import coinstac
        coinstac.start(None, compute)
    except ImportError:
        raise SystemExit('run under COINSTAC, or use examples/run_simulator.py')
