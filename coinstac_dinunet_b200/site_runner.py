"""``SiteRunner`` - run ONE site offline over a COINSTAC-simulator folder layout.

Parity: coinstac_dinunet/site_runner.py:8-45 (including the historic ``taks_id`` spelling of
the first parameter).  Layout expected under ``data_path``::

    inputspec.json                      list with one {"key": {"value": v}} dict per site
    input/local<i>/simulatorRun/...     the site's data (-> state['baseDirectory'])
    output/local<i>/simulatorRun/       created; run artefacts go to ``_srun_<task>``
"""
import json as _json
import os as _os

from .config.keys import Phase
from .distrib.nodes.local import COINNLocal


class SiteRunner(COINNLocal):
    def __init__(self, taks_id, data_path='test', site_index=0, **kw):
        with open(_os.path.join(data_path, 'inputspec.json')) as fp:
            spec = _json.loads(fp.read())[site_index]
        cache = {k: v['value'] for k, v in spec.items()}

        site = f'local{site_index}'
        out_dir = _os.path.join(data_path, 'output', site, 'simulatorRun', f'_srun_{taks_id}')
        state = {
            'baseDirectory': _os.path.join(data_path, 'input', site, 'simulatorRun'),
            'outputDirectory': out_dir,
            'transferDirectory': out_dir,
            'clientId': site,
        }
        _os.makedirs(out_dir, exist_ok=True)
        super().__init__(task_id=taks_id, cache=cache, input={}, state=state, **kw)

    def _round(self, extra_input, trainer_cls, dataset_cls, datahandle_cls):
        self.input = {**self.input, **extra_input}
        self.out = {}
        self.compute(None, trainer_cls, dataset_cls, datahandle_cls)
        return self.out

    def run(self, trainer_cls, dataset_cls, datahandle_cls, **kw):
        """init_runs, then next_run with this site flagged as the pre-training site so the
        whole ``train_local`` loop executes; finishes in ``phase = pre_computation``."""
        self.cache.update(**kw)
        self._round({'phase': Phase.INIT_RUNS}, trainer_cls, dataset_cls, datahandle_cls)

        self.cache['verbose'] = True
        self._pretrain_args = {k: v for k, v in self.cache.items()
                               if k not in ('nn', 'optimizer', 'device', 'dataset')}
        runs = {self.state['clientId']: {'split_ix': '0', 'seed': 1, 'pretrain': True}}
        return self._round({'phase': Phase.NEXT_RUN, 'global_runs': runs},
                           trainer_cls, dataset_cls, datahandle_cls)
