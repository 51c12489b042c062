import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on a B200 via gpurun)')
    config.addinivalue_line('markers', 'multigpu: needs >= 2 CUDA devices')


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        have = torch.cuda.is_available()
        ndev = torch.cuda.device_count() if have else 0
    except Exception:
        have, ndev = False, 0
    for item in items:
        if 'gpu' in item.keywords and not have:
            item.add_marker(pytest.mark.skip(reason='no CUDA device'))
        if 'multigpu' in item.keywords and ndev < 2:
            item.add_marker(pytest.mark.skip(reason='needs >= 2 GPUs'))


@pytest.fixture
def fs_sites(tmp_path):
    """Two synthetic FreeSurfer sites laid out like the COINSTAC simulator + an engine."""
    from coinstac_dinunet_b200.engine import InProcessEngine
    from coinstac_dinunet_b200.models import write_synthetic_site

    def make(n_sites=2, sizes=(24, 18), spec=None, **kw):
        base = dict(task_id='fsv', mode='train', data_dir='data', labels_file='labels.json', input_size=66,
                    num_class=2, batch_size=4, epochs=2, num_folds=3, learning_rate=1e-2, seed=7,
                    monitor_metric='f1', metric_direction='maximize', log_header='Loss|Accuracy,F1',
                    verbose=False)
        base.update(spec or {})
        eng = InProcessEngine(tmp_path / 'work', n_sites=n_sites, inputspec=base, **kw)
        for i, site in enumerate(eng.site_ids):
            write_synthetic_site(eng.site_state[site]['baseDirectory'], sizes[i % len(sizes)], (66,), seed=i)
        return eng
    return make
