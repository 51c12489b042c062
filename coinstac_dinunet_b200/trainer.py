"""``COINNTrainer`` - NNTrainer whose live state persists in the node ``cache`` between
engine rounds, plus the distributed validation / test passes.

Parity: coinstac_dinunet/trainer.py:15-80.
"""
from abc import ABC
from os import sep as _sep

from . import config as _conf
from . import metrics as _metrics
from .config.keys import Key
from .nn import NNTrainer as _NNTrainer
from .utils.utils import performance_improved_


class COINNTrainer(_NNTrainer, ABC):
    def __init__(self, **kw):
        super().__init__(**kw)
        # The node object is rebuilt every round; models/optimizers/devices survive in cache.
        self.nn = self.cache.setdefault('nn', {})
        self.device = self.cache.setdefault('device', {})
        self.optimizer = self.cache.setdefault('optimizer', {})

    def _save_if_better(self, epoch, val_metrics):
        """Pre-training: snapshot ``weights.tar`` into the transfer dir on improvement."""
        out = {}
        score = val_metrics.extract(self.cache['monitor_metric'])
        if performance_improved_(epoch, score, self.cache):
            out['weights_file'] = _conf.weights_file
            self.save_checkpoint(file_path=self.state['transferDirectory'] + _sep + out['weights_file'])
        return out

    @staticmethod
    def _as_list(ds):
        if ds and not isinstance(ds, list):
            return [ds]
        return ds

    def validation_distributed(self, dataset_cls):
        out = {}
        val = self._as_list(self.data_handle.dataset.get('validation'))
        if val:
            avg, met = self.evaluation(mode='validation', save_pred=False, dataset_list=val,
                                       use_padded_sampler=True)
            out[Key.VALIDATION_SERIALIZABLE] = [{'averages': avg.serialize(), 'metrics': met.serialize()}]
        self.cache['cursor'] = 0
        return out

    def test_distributed(self, dataset_cls):
        import os as _os
        out = {}
        best = self.cache['log_dir'] + _sep + self.cache['best_nn_state']
        if _os.path.exists(best):  # quirk §8.5-11: the reference crashes when no best exists
            self.load_checkpoint(best)
        test = self._as_list(self.data_handle.get_test_dataset(dataset_cls))
        if test:
            avg, met = self.evaluation(mode='test', save_pred=True, dataset_list=test)
            out[Key.TEST_SERIALIZABLE] = [{'averages': avg.serialize(), 'metrics': met.serialize()}]
        return out

    def set_monitor_metric(self):
        """Must be set from COINNLocal's constructor"""

    def set_log_headers(self):
        """Must be set from COINNLocal's constructor"""

    def new_metrics(self):
        """Binary -> Prf1a (or AUC when monitored), multi-class -> ConfusionMatrix."""
        n = self.cache.get('num_class')
        if n == 2:
            monitored = self.cache.get('monitor_metric')
            if monitored == 'auc':
                return _metrics.AUCROCMetrics()
            if monitored in ('precision', 'recall', 'accuracy', 'overlap', 'f1'):
                return _metrics.Prf1a()
        elif n is not None and n > 2:
            return _metrics.ConfusionMatrix(num_classes=n)
        return _metrics.Prf1a() if n == 2 else _metrics.COINNMetrics()
