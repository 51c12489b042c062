"""``COINNRemote`` - the aggregator state machine: global barriers, fold/epoch bookkeeping,
model selection, score aggregation and the final results archive.

Protocol parity: coinstac_dinunet/distrib/nodes/remote.py:22-310 and SURVEY §3.0/§3.2.
Every transition is gated on *all* sites agreeing (``check(all, key, value, input)``), which is
the protocol's global barrier.
"""
import datetime as _datetime
import json as _json
import os as _os
import shutil as _shutil
import traceback as _tback

from ... import config as _conf
from ... import utils as _utils
from ...config.keys import AGG_Engine, Key, Mode, Phase
from ...utils import performance_improved_, stop_training_
from ...utils.logger import lazy_debug
from ...vision import plotter as _plot
from ..reducer import COINNReducer as _dSGDReducer


def _engine_reducers():
    from ..powersgd import PowerSGDReducer
    from ..rankdad import DADReducer
    return {AGG_Engine.dSGD: _dSGDReducer, AGG_Engine.rankDAD: DADReducer, AGG_Engine.powerSGD: PowerSGDReducer}


class EmptyDataHandle:
    """Data-less stand-in so a trainer (for ``new_metrics``/``new_averages``) exists remotely."""

    def __init__(self, cache, input, state):
        self.cache, self.input, self.state = cache, input, state


def _gather(keys, data, mode='append'):
    """Collect ``keys`` from an iterable of dicts. ``append`` keeps one entry per dict,
    ``extend`` concatenates list-valued entries."""
    if mode not in ('append', 'extend'):
        raise AssertionError(f"Invalid mode:{mode}. Has to be ['append', 'extend']")
    rows = list(data)
    res = {}
    for k in keys:
        bucket = []
        for row in rows:
            val = row.get(k)
            if not val:
                continue
            if mode == 'append':
                bucket.append(val)
            else:
                bucket = bucket + list(val)
        res[k] = bucket
    return res


def check(logic, k, v, kw):
    """Barrier predicate: ``logic`` (``all``/``any``) over ``site[k] == v`` for every site."""
    return logic([site_vars.get(k) == v for site_vars in kw.values()])


class _AggRound:
    __slots__ = ('mp_pool', 'reducer_cls', 'trainer', 'reducer')

    def __init__(self, mp_pool, reducer_cls, trainer):
        self.mp_pool, self.reducer_cls, self.trainer, self.reducer = mp_pool, reducer_cls, trainer, None


class COINNRemote:
    def __init__(self, cache: dict = None, input: dict = None, state: dict = None, verbose=False, **kw):
        self.out = {}
        self.cache = cache
        self.cache.update(**kw)
        self.input = _utils.FrozenDict(input)
        self.state = _utils.FrozenDict(state)
        self.cache['verbose'] = verbose
        if not self.cache.get(Key.ARGS_CACHED):
            first_site = next(iter(self.input.values()))
            self.cache.update(**first_site['shared_args'])
            for k in ('resume', 'checkpoint_epochs'):            # (ours) aggregator-side switches: a constructor kwarg wins
                if kw.get(k):
                    self.cache[k] = kw[k]
            self.cache[Key.ARGS_CACHED] = True

    # ------------------------------------------------------------------- folds
    def _init_runs(self):
        self.cache['seed'] = self.cache.setdefault('seed', _conf.current_seed)
        self.cache[Key.GLOBAL_TEST_SERIALIZABLE] = []
        self.cache['data_size'] = {site: sv.get('data_size') for site, sv in self.input.items()}
        # a stack: fold 0 is popped first
        self.cache['folds'] = [{'split_ix': str(f), 'seed': self.cache['seed']}
                               for f in reversed(range(self.cache['num_folds']))]
        self._maybe_resume()

    # ------------------------------------------------------------------ resume
    def _resume_file(self):
        return _os.path.join(self.state['outputDirectory'], str(self.cache['task_id']), 'resume.json')

    def _maybe_resume(self):
        """Resume (the reference has none): with ``resume=True`` folds recorded as finished by an earlier
        (crashed / stopped) run over the same output directory are skipped and their test scores are carried into the
        final cross-fold aggregate; if that run had committed an epoch-level resume point for the fold it was in
        (``checkpoint_epochs``, see ``_request_resume_point``), that fold continues from there instead of epoch 0."""
        path = self._resume_file()
        if not self.cache.get('resume') or not _os.path.exists(path):
            return
        with open(path) as fp:
            done = _json.load(fp)
        if done.get('seed') is not None:
            self.cache['seed'] = done['seed']
        finished = set(done.get('completed_folds', []))
        self.cache['folds'] = [f for f in self.cache['folds'] if f['split_ix'] not in finished]
        for f in self.cache['folds']:
            f['seed'] = self.cache['seed']
        self.cache[Key.GLOBAL_TEST_SERIALIZABLE] = list(done.get('global_test_serializable', []))
        self.cache['resumed_folds'] = sorted(finished)
        self.cache['_resume_point'] = done.get('in_progress')

    # Epoch-level resume points are a two-phase commit, because the sites' checkpoints and the aggregator's counters must
    # describe the same epoch: (1) after a validation round the aggregator asks for a resume point (``save_resume_point =
    # epoch``) and stashes its own state of that moment; (2) every site writes ``resume.<task>-<fold>.e<epoch>.pt`` and
    # answers ``resume_point_saved = epoch``; (3) only when ALL sites answered does the aggregator record the point in
    # ``resume.json`` and announce ``resume_point_committed`` (sites then prune older files).  A crash anywhere in between
    # leaves the previous committed point - and the files it names - intact.
    def _request_resume_point(self, next_mode):
        every = int(self.cache.get('checkpoint_epochs') or 0)
        if not every or next_mode != Mode.TRAIN:
            return
        validations = self.cache['epoch'] // max(int(self.cache['validation_epochs']), 1)
        if validations % every:
            return
        self.out['save_resume_point'] = int(self.cache['epoch'])
        self.cache['_pending_resume'] = {
            'split_ix': self.cache['fold']['split_ix'], 'epoch': int(self.cache['epoch']),
            'best_val_epoch': self.cache['best_val_epoch'], 'best_val_score': self.cache['best_val_score'],
            'train_log': [list(r) for r in self.cache[Key.TRAIN_LOG]],
            'validation_log': [list(r) for r in self.cache[Key.VALIDATION_LOG]]}

    def _commit_resume_point(self, rt=None):
        pend = self.cache.get('_pending_resume')
        if not pend or not check(all, 'resume_point_saved', pend['epoch'], self.input):
            return
        path = self._resume_file()
        _os.makedirs(_os.path.dirname(path), exist_ok=True)
        rec = {'completed_folds': [], 'seed': self.cache.get('seed'),
               'global_test_serializable': self.cache[Key.GLOBAL_TEST_SERIALIZABLE]}
        if _os.path.exists(path):
            with open(path) as fp:
                rec.update(_json.load(fp))
        rec['in_progress'] = pend
        tmp = path + '.tmp'
        with open(tmp, 'w') as fp:
            _json.dump(rec, fp)
        _os.replace(tmp, path)                       # atomic: a crash never leaves half a record
        self.out['resume_point_committed'] = pend['epoch']
        self.cache.pop('_pending_resume', None)

    def _record_fold_done(self):
        path = self._resume_file()
        _os.makedirs(_os.path.dirname(path), exist_ok=True)
        prev = {'completed_folds': []}
        if _os.path.exists(path):
            with open(path) as fp:
                prev = _json.load(fp)
        folds = list(dict.fromkeys(list(prev.get('completed_folds', [])) + [self.cache['fold']['split_ix']]))
        with open(path, 'w') as fp:
            _json.dump({'completed_folds': folds, 'seed': self.cache.get('seed'),
                        'global_test_serializable': self.cache[Key.GLOBAL_TEST_SERIALIZABLE]}, fp)

    def _next_run(self, trainer):
        """Pop the next fold, reset epoch/score/log state, tell each site its fold, seed and
        whether it is the (single) pre-training site - the one with most training data."""
        fold = self.cache['fold'] = self.cache['folds'].pop()
        self.cache['log_dir'] = _os.path.join(self.state['outputDirectory'], self.cache['task_id'],
                                              f"fold_{fold['split_ix']}")
        _os.makedirs(self.cache['log_dir'], exist_ok=True)
        trainer.init_nn(set_devices=True)

        maximize = self.cache['metric_direction'] == 'maximize'
        self.cache.update(epoch=0, best_val_epoch=0, best_val_score=0 if maximize else _conf.max_size)
        for k in (Key.TRAIN_LOG, Key.VALIDATION_LOG, Key.TEST_METRICS):
            self.cache[k] = []

        self.cache.pop('_pending_resume', None)
        point = self.cache.pop('_resume_point', None)
        resume_epoch = None
        if point and str(point.get('split_ix')) == fold['split_ix']:
            resume_epoch = int(point['epoch'])
            self.cache.update(epoch=resume_epoch, best_val_epoch=point['best_val_epoch'],
                              best_val_score=point['best_val_score'], resumed_epoch=resume_epoch)
            self.cache[Key.TRAIN_LOG] = [list(r) for r in point.get('train_log', [])]
            self.cache[Key.VALIDATION_LOG] = [list(r) for r in point.get('validation_log', [])]

        train_sizes = {site: self.cache['data_size'][site][fold['split_ix']].get('train', 0)
                       for site in self.input}
        biggest = max(train_sizes, key=train_sizes.get)
        return {site: {**fold, 'pretrain': site == biggest and resume_epoch is None, 'resume_epoch': resume_epoch}
                for site in self.input}

    # ------------------------------------------------------------------ scores
    def _reduce_scores(self, trainer, entries):
        parts = _gather(['averages', 'metrics'], entries, 'append')
        averages, metrics = trainer.new_averages(), trainer.new_metrics()
        averages.reduce_sites(parts['averages'])
        metrics.reduce_sites(parts['metrics'])
        return averages, metrics

    def _accumulate_epoch_info(self, trainer):
        train = _gather([Key.TRAIN_SERIALIZABLE], self.input.values(), 'extend')[Key.TRAIN_SERIALIZABLE]
        val = _gather([Key.VALIDATION_SERIALIZABLE], self.input.values(), 'extend')[Key.VALIDATION_SERIALIZABLE]
        out = {}
        out['train_averages'], out['train_metrics'] = self._reduce_scores(trainer, train)
        out['val_averages'], out['val_metrics'] = self._reduce_scores(trainer, val)
        return out

    def _on_epoch_end(self, reducer):
        info = self._accumulate_epoch_info(reducer.trainer)
        self.cache[Key.TRAIN_LOG].append([*info['train_averages'].get(), *info['train_metrics'].get()])
        self._save_if_better(**info)
        if info.get('val_averages'):
            self.cache[Key.VALIDATION_LOG].append([*info['val_averages'].get(), *info['val_metrics'].get()])
        if lazy_debug(self.cache['epoch']):
            _plot.plot_progress(self.cache, self.cache['log_dir'], plot_keys=[Key.TRAIN_LOG, Key.VALIDATION_LOG])
        return info

    def _on_run_end(self, trainer):
        """A fold finished (every site is ``next_run_waiting``): fold test scores -> CSV/plots/logs."""
        entries = _gather([Key.TEST_SERIALIZABLE], self.input.values(), 'extend')[Key.TEST_SERIALIZABLE]
        averages, metrics = self._reduce_scores(trainer, entries)
        self.cache[Key.TEST_METRICS].append([*averages.get(), *metrics.get()])
        self.cache[Key.GLOBAL_TEST_SERIALIZABLE].append(
            {'averages': averages.serialize(), 'metrics': metrics.serialize()})

        _plot.plot_progress(self.cache, self.cache['log_dir'], plot_keys=[Key.TRAIN_LOG, Key.VALIDATION_LOG])
        _utils.save_scores(self.cache, self.cache['log_dir'], file_keys=[Key.TEST_METRICS])
        snapshot = {**self.cache}
        snapshot[Key.GLOBAL_TEST_SERIALIZABLE] = snapshot[Key.GLOBAL_TEST_SERIALIZABLE][-1]
        _utils.save_cache(snapshot, self.cache['log_dir'])
        self._record_fold_done()

    def _send_global_scores(self, trainer):
        """All folds done: cross-fold score, ``global_test_metrics.csv`` and the results zip."""
        averages, metrics = self._reduce_scores(trainer, self.cache[Key.GLOBAL_TEST_SERIALIZABLE])
        self.cache[Key.GLOBAL_TEST_METRICS] = [[*averages.get(), *metrics.get()]]
        task_dir = self.state['outputDirectory'] + _os.sep + self.cache['task_id']
        _utils.save_scores(self.cache, task_dir, file_keys=[Key.GLOBAL_TEST_METRICS])

        stamp = '_'.join(str(_datetime.datetime.now()).split(' '))
        out = {'results_zip': f"{self.cache['task_id']}_{self.cache['agg_engine']}_{stamp}"}
        _shutil.make_archive(f"{self.state['transferDirectory']}{_os.sep}{out['results_zip']}", 'zip', task_dir)
        return out

    # -------------------------------------------------------------------- modes
    def _set_mode(self, mode=None):
        return {site: (mode if mode else sv.get('mode', 'N/A')) for site, sv in self.input.items()}

    def _pre_compute(self):
        """Broadcast the pre-trained checkpoint: first site that shipped ``weights_file`` wins."""
        out = {}
        for site, sv in self.input.items():
            if sv.get('weights_file') is not None:
                src = self.state['baseDirectory'] + _os.sep + site + _os.sep + sv['weights_file']
                out['pretrained_weights'] = f'pretrained_{_conf.weights_file}'
                out['pretrained_site'] = site            # device transports broadcast GPU-to-GPU from this site (C5)
                _shutil.copy(src, self.state['transferDirectory'] + _os.sep + out['pretrained_weights'])
                break
        return out

    # ------------------------------------------------------------------ compute
    # The aggregator is a barrier table: every row is (key, value, handler); a row fires when ALL sites report
    # ``site[key] == value`` (the protocol's global barrier, ``check``).  Rows are evaluated top to bottom within a
    # round; the computation rows are nested under the row that recognises the computation phase.
    def _on_all_init(self, rt):
        self._init_runs()
        self._start_next_fold_or_finish(rt)

    def _on_all_pre_computation(self, rt):
        self.out.update(**self._pre_compute())
        self.out['phase'] = Phase.PRE_COMPUTATION

    def _on_all_computation(self, rt):
        rt.reducer = self._get_reducer_cls(rt.reducer_cls)(trainer=rt.trainer, mp_pool=rt.mp_pool)
        self.out['phase'] = Phase.COMPUTATION
        self._fire(self._COMPUTATION_BARRIERS, rt)

    def _on_all_reduce(self, rt):
        self.out.update(**rt.reducer.reduce())

    def _on_all_validation_waiting(self, rt):
        self.cache['epoch'] += 1
        validate = self.cache['epoch'] % self.cache['validation_epochs'] == 0
        self.out['global_modes'] = self._set_mode(mode=Mode.VALIDATION if validate else Mode.TRAIN)

    def _on_all_train_waiting(self, rt):
        info = self._on_epoch_end(rt.reducer)
        nxt = self._next_epoch(**info)['mode']
        self.out['global_modes'] = self._set_mode(mode=nxt)
        self._request_resume_point(nxt)

    def _on_all_next_run_waiting(self, rt):
        self._on_run_end(rt.trainer)
        self._start_next_fold_or_finish(rt)

    def _start_next_fold_or_finish(self, rt):
        if len(self.cache['folds']) > 0:
            self.out['global_runs'] = self._next_run(rt.trainer)
            self.out['phase'] = Phase.NEXT_RUN
        else:                                       # also: a resumed run with nothing left to train
            self.out.update(**self._send_global_scores(rt.trainer))
            self.out['phase'] = Phase.SUCCESS

    def _echo_modes(self, rt):
        self.out['global_modes'] = self._set_mode()

    _BARRIERS = (
        ('phase', Phase.INIT_RUNS, '_on_all_init'),
        ('phase', Phase.PRE_COMPUTATION, '_on_all_pre_computation'),
        (None, None, '_echo_modes'),                                   # unconditional: default mode echo
        (None, None, '_commit_resume_point'),                          # unconditional: phase 3 of a resume point, if one is pending
        ('phase', Phase.COMPUTATION, '_on_all_computation'),
        ('phase', Phase.NEXT_RUN_WAITING, '_on_all_next_run_waiting'),
    )
    _COMPUTATION_BARRIERS = (
        ('reduce', True, '_on_all_reduce'),
        ('mode', Mode.VALIDATION_WAITING, '_on_all_validation_waiting'),
        ('mode', Mode.TRAIN_WAITING, '_on_all_train_waiting'),
    )

    def _fire(self, table, rt):
        for key, value, handler in table:
            if key is None or check(all, key, value, self.input):
                getattr(self, handler)(rt)

    def compute(self, mp_pool, trainer_cls, reducer_cls: callable = _dSGDReducer, **kw):
        rt = _AggRound(mp_pool, reducer_cls,
                       trainer_cls(data_handle=EmptyDataHandle(cache=self.cache, input=self.input, state=self.state)))
        self.out['phase'] = self.input.get('phase', Phase.INIT_RUNS)
        self._fire(self._BARRIERS, rt)

    def _next_epoch(self, **kw):
        # NB ``>`` (not ``>=``): the reference trains epochs+1 epochs (quirk §8.5-5) - kept for parity
        done = self.cache['epoch'] > self.cache['epochs']
        return {'mode': Mode.TEST if (done or self._stop_early(**kw)) else Mode.TRAIN}

    def _save_if_better(self, **kw):
        if kw.get('val_metrics'):
            score = kw['val_metrics'].extract(self.cache['monitor_metric'])
            self.out['save_current_as_best'] = performance_improved_(self.cache['epoch'], score, self.cache)

    def _stop_early(self, **kw):
        return stop_training_(self.cache['epoch'], self.cache)

    def _get_reducer_cls(self, reducer_cls):
        return _engine_reducers().get(self.cache.get('agg_engine'), reducer_cls)

    def __call__(self, *args, **kwargs):
        try:
            self.compute(*args, **kwargs)
            return {'output': self.out, 'success': check(all, 'phase', Phase.SUCCESS, self.input)}
        except Exception:
            _tback.print_exc()
            raise Exception(self.out)
