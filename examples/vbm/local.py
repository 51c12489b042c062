"""Site entry point of the VBM 3-D CNN computation (COINSTAC local node)."""
from multiprocessing.pool import ThreadPool

from coinstac_dinunet_b200 import COINNLocal
from coinstac_dinunet_b200.models import VBMDataset, VBMTrainer

_cache, _pool = {}, None          # the node process is long-lived: live objects persist in `_cache`


def compute(args):
    global _pool
    _pool = _pool or ThreadPool(2)
    cache = args.get('cache') if args.get('cache') is not None else _cache
    node = COINNLocal(cache=cache, input=args['input'], state=args['state'], task_id="vbm", epochs=31, batch_size=8,
                      learning_rate=1e-3, monitor_metric='f1', log_header='Loss|Accuracy,F1,Precision,Recall')
    return node(_pool, VBMTrainer, VBMDataset)


if __name__ == '__main__':
    try:
        import coinstac
        coinstac.start(compute, None)
    except ImportError:
        raise SystemExit('run under COINSTAC, or use examples/run_simulator.py')
