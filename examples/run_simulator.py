#!/usr/bin/env python
"""Run an example computation in the in-process engine on synthetic sites:  python examples/run_simulator.py fsv|vbm|custom"""
import importlib.util
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from coinstac_dinunet_b200.engine import InProcessEngine  # noqa: E402
from coinstac_dinunet_b200.models import write_synthetic_site  # noqa: E402


def _load(path, name):
    here = os.path.dirname(path)
    if here not in sys.path:                 # entry scripts may import their siblings (`from local import MyTrainer`)
        sys.path.insert(0, here)
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main(which='fsv', n_sites=2):
    here = os.path.join(ROOT, 'examples', which)
    sys.modules.pop('local', None)          # a sibling import of another example must not be reused
    local, remote = _load(os.path.join(here, 'local.py'), f'{which}_local'), _load(os.path.join(here, 'remote.py'), f'{which}_remote')
    shape = (66,) if which == 'fsv' else (1, 33, 37, 33)
    spec = dict(mode='train', data_dir='data', labels_file='labels.json', num_class=2, split_ratio=[0.6, 0.2, 0.2],
                epochs=3, batch_size=4, learning_rate=1e-2, input_size=66, input_shape=list(shape), seed=3)
    work = tempfile.mkdtemp(prefix=f'coinn_{which}_')
    eng = InProcessEngine(work, n_sites=n_sites, inputspec=spec)
    for i, site in enumerate(eng.site_ids):
        write_synthetic_site(eng.site_state[site]['baseDirectory'], 24 + 6 * i, shape, seed=i)
    rounds = eng.run(lambda site, cache, inp, state: local.compute({'cache': cache, 'input': inp, 'state': state}),
                     lambda cache, inp, state: remote.compute({'cache': cache, 'input': inp, 'state': state}))
    out = os.path.join(eng.remote_state['outputDirectory'], which)
    print(f'{which}: {rounds} rounds, results in {out}:', sorted(os.listdir(out)))
    return eng


if __name__ == '__main__':
    main(sys.argv[1] if len(sys.argv) > 1 else 'fsv')
