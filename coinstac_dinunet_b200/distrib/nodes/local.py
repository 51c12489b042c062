"""``COINNLocal`` - the site-side, re-entrant state machine driven once per engine round.

Protocol parity: coinstac_dinunet/distrib/nodes/local.py:29-295 and the round table in
SURVEY §3.0 (phases, JSON keys, file names, artefact layout).  The class is stateless between
rounds - everything persistent lives in ``cache`` (quirk §8.5-17).

Transport: with ``cache['transport'] == 'file'`` (default) one round == one optimizer step,
exactly like the reference.  With ``'nvlink'``/``'nccl'`` the learner runs *a whole local epoch*
of fused steps per round (gradient exchange inside the kernel) and the JSON control plane is
touched once per epoch.
"""
import json as _json
import os as _os
import shutil as _shutil
import time as _time
import traceback as _tback
from os import sep as _sep
from typing import List as _List

from ... import config as _conf
from ... import utils as _utils
from ...config.keys import AGG_Engine, Key, Mode, Phase, Transport
from ...data import COINNDataHandle as _DataHandle
from ...utils import FrozenDict as _FrozenDict
from ..learner import COINNLearner as _dSGDLearner


def _engine_learners():
    from ..powersgd import PowerSGDLearner
    from ..rankdad import DADLearner
    return {AGG_Engine.dSGD: _dSGDLearner, AGG_Engine.rankDAD: DADLearner, AGG_Engine.powerSGD: PowerSGDLearner}


class _Round:
    """What one engine round carries between handlers (the node object itself is rebuilt every round)."""
    __slots__ = ('mp_pool', 'trainer_cls', 'dataset_cls', 'datahandle_cls', 'learner_cls', 'trainer', 'learner', 'modes')

    def __init__(self, **kw):
        self.trainer = self.learner = None
        self.modes = []
        for k, v in kw.items():
            setattr(self, k, v)


class COINNLocal:
    _PROMPT_TASK_ = "Task id must be given."
    _PROMPT_MODE_ = f"Mode must be provided and should be one of {[Mode.TRAIN.value, Mode.TEST.value]}."

    def __init__(self, cache: dict = None, input: dict = None, state: dict = None,
                 task_id='nn_task', mode: str = None, batch_size: int = 8, local_iterations: int = 1,
                 epochs: int = 31, validation_epochs: int = 1, learning_rate: float = 0.001,
                 gpus: _List[int] = None, pin_memory: bool = False, num_workers: int = 0,
                 load_limit: int = _conf.max_size, load_sparse=False, pretrained_path: str = None,
                 patience: int = None, num_folds: int = None, split_ratio=None,
                 pretrain_args: dict = None, dataloader_args: dict = None, verbose=False,
                 monitor_metric='f1', metric_direction='maximize', log_header='Loss|Accuracy,F1',
                 agg_engine='dSGD', num_reducers=2, precision_bits=32, checkpoint_epochs: int = 0,
                 resume: bool = False, **kw):
        self.out = {}
        self.cache = cache
        self.input = _FrozenDict(input)
        self.state = _FrozenDict(state)

        defaults = dict(
            task_id=task_id, mode=mode, batch_size=batch_size, local_iterations=local_iterations,
            epochs=epochs, validation_epochs=validation_epochs, learning_rate=learning_rate, gpus=gpus,
            pin_memory=pin_memory, num_workers=num_workers, load_limit=load_limit, load_sparse=load_sparse,
            pretrained_path=pretrained_path, patience=patience if patience else epochs,
            split_ratio=split_ratio, num_folds=num_folds, verbose=verbose, monitor_metric=monitor_metric,
            metric_direction=metric_direction, log_header=log_header, agg_engine=agg_engine,
            num_reducers=num_reducers, precision_bits=precision_bits,
            checkpoint_epochs=checkpoint_epochs, resume=resume)      # (ours) epoch-level resume points, see COINNRemote
        defaults.update(**kw)
        self._args = _FrozenDict(defaults)
        self._pretrain_args = pretrain_args if pretrain_args else {}
        self._dataloader_args = dataloader_args if dataloader_args else {}

        if not self.cache.get(Key.ARGS_CACHED):
            self._cache_args_once()

    def _cache_args_once(self):
        """Merge configuration into ``cache`` exactly once, highest priority first:
        ``input`` → ``input['<task>_args']`` → ``input['<engine>_args']`` →
        ``input['<task>_data_conf']`` (only keys the two arg groups did not set) →
        constructor defaults for whatever is still ``None`` (ref local.py:93-117)."""
        inp = self.input
        self.cache.update(**inp)
        task_args = inp.get(f"{inp.get('task_id')}_args", {})
        engine_args = inp.get(f"{inp.get('agg_engine')}_args", {})
        self.cache.update(**task_args)
        self.cache.update(**engine_args)
        for k, v in inp.get(f"{inp.get('task_id')}_data_conf", {}).items():
            if k not in task_args and k not in engine_args:
                self.cache[k] = v
        for k, v in self._args.items():
            if self.cache.get(k) is None:
                self.cache[k] = v

        assert self.cache['task_id'] is not None, self._PROMPT_TASK_
        assert self.cache['mode'] in (Mode.TRAIN, Mode.TEST), self._PROMPT_MODE_
        if self.cache['mode'] == Mode.TRAIN:
            assert self.cache['split_ratio'] or self.cache['num_folds'], \
                "Split ratio or K(num k-folds) is needed."
        self.cache[Key.ARGS_CACHED] = True

    # ------------------------------------------------------------ phase handlers
    def _init_runs(self, trainer):
        out = {}
        out.update(trainer.data_handle.prepare_data())
        self.cache['num_folds'] = len(self.cache['splits'])
        trainer.init_nn(set_devices=True)

        out['data_size'] = {}
        for fold, name in self.cache['splits'].items():
            with open(self.cache['split_dir'] + _sep + name) as fp:
                split = _json.loads(fp.read())
            out['data_size'][fold] = {part: len(files) for part, files in split.items()}
        return out

    def _next_run(self, trainer):
        self.cache.update(cursor=0)
        self.cache[Key.TRAIN_SERIALIZABLE] = []
        ix = self.cache['split_ix']
        self.cache['split_file'] = self.cache['splits'][ix]
        self.cache['log_dir'] = _os.path.join(self.state['outputDirectory'], self.cache['task_id'], f"fold_{ix}")
        _os.makedirs(self.cache['log_dir'], exist_ok=True)

        # a new fold starts from scratch: drop engine state that is tied to the old modules
        for stale in ('_arena', '_graph_step', 'powerSGD_state', 'local_epoch'):
            self.cache.pop(stale, None)
        trainer.init_nn(init_model=True, init_optim=True, set_devices=True, init_weights=True)
        self.cache['best_nn_state'] = f"best.{self.cache['task_id']}-{ix}.pt"
        self.cache['latest_nn_state'] = f"latest.{self.cache['task_id']}-{ix}.pt"
        out = {'phase': Phase.COMPUTATION}
        if self.cache.get('resume_epoch') is not None:           # the aggregator found a committed resume point
            out['resumed_epoch'] = self._load_resume_point(trainer, int(self.cache['resume_epoch']))
        return out

    # ------------------------------------------------------- epoch-level resume points (see COINNRemote._request_resume_point)
    def _resume_path(self, epoch, ext='pt'):
        return _os.path.join(self.cache['log_dir'], f"resume.{self.cache['task_id']}-{self.cache['split_ix']}.e{int(epoch)}.{ext}")

    def _load_resume_point(self, trainer, epoch):
        path = self._resume_path(epoch)
        if not _os.path.exists(path):
            raise FileNotFoundError(f"{self.state['clientId']}: the aggregator resumes fold {self.cache['split_ix']} at epoch "
                                    f"{epoch} but {path} does not exist (was the output directory replaced?)")
        trainer.load_checkpoint(file_path=path)
        with open(self._resume_path(epoch, 'json')) as fp:
            meta = _json.load(fp)
        self.cache['local_epoch'] = int(meta.get('local_epoch', 0))
        extra = self._resume_path(epoch, 'engine.pt')
        if _os.path.exists(extra):                               # engine state that lives outside model / optimizer
            import torch as _torch
            self.cache.update(_torch.load(extra, weights_only=False))
        self.cache['resume_committed'] = epoch
        return epoch

    def _pretrain_local(self, trainer_cls, datahandle_cls, train_dataset, validation_dataset):
        """Optional single-site warm-up on the site holding the most data (chosen by the
        remote via ``global_runs[site]['pretrain']``); its best weights are then broadcast."""
        out = {'phase': Phase.COMPUTATION}
        wants = self._pretrain_args.get('epochs', 0) > 0
        if wants and self.cache.get('pretrain'):
            overrides = dict(self.cache.get('pretrain_args') or self._pretrain_args)
            saved = {k: self.cache.get(k) for k in overrides}
            self.cache.update(overrides)  # quirk §8.5-4 fixed: pretrain args really apply
            try:
                trainer = trainer_cls(data_handle=datahandle_cls(
                    cache=self.cache, input=self.input, state=self.state, dataloader_args=self._dataloader_args))
                trainer.init_nn()
                trainer.init_training_cache()
                out.update(**trainer.train_local(train_dataset, validation_dataset))
            finally:
                for k, v in saved.items():
                    if v is None:
                        self.cache.pop(k, None)
                    else:
                        self.cache[k] = v
            out['phase'] = Phase.PRE_COMPUTATION
        if wants and any(r.get('pretrain') for r in self.input['global_runs'].values()):
            out['phase'] = Phase.PRE_COMPUTATION
        return out

    # ------------------------------------------------------------------ compute
    # The site is a table-driven machine.  One round = (1) the ENTRY handler of the phase the aggregator sent,
    # (2) learner construction + mode echo, (3) the EXIT handler of the phase the entry handler left in ``out``.
    # The computation phase is itself a rule table over (input flags, global modes): every rule whose guard holds
    # fires, in order.  Adding a phase / a rule is a table entry, not another branch.
    def _enter_init_runs(self, rt):
        self.out.update(**self._init_runs(rt.trainer))
        # share the constructor-level arguments with the aggregator and freeze them
        shared = {k: self.cache[k] for k in self._args}
        # also share what the aggregator needs to build the same metric objects
        for k in ('num_class', *self.cache.get('shared_keys', ())):
            if k in self.cache and k not in shared:
                shared[k] = self.cache[k]
        self.cache['frozen_args'] = _FrozenDict(shared)
        self.out['shared_args'] = self.cache['frozen_args']

    def _enter_next_run(self, rt):
        self.cache.update(**self.input['global_runs'][self.state['clientId']])
        self.out.update(**self._next_run(rt.trainer))
        if self.cache['mode'] == Mode.TRAIN:
            dh = rt.trainer.data_handle
            self.out.update(**self._pretrain_local(rt.trainer_cls, rt.datahandle_cls,
                                                   dh.get_train_dataset(rt.dataset_cls),
                                                   dh.get_validation_dataset(rt.dataset_cls)))

    def _enter_pre_computation(self, rt):
        if not self.input.get('pretrained_weights'):
            return
        path = self.state['baseDirectory'] + _sep + self.input['pretrained_weights']
        if not self._device_broadcast(rt, path):
            rt.trainer.load_checkpoint(file_path=path)
        self.out['phase'] = Phase.COMPUTATION

    def _device_broadcast(self, rt, path):
        """C5 on the device transports: only the pre-training site reads its checkpoint; parameters, buffers and
        optimizer moments then go GPU-to-GPU over NVLink peer copies (``DistArena.broadcast_from``) instead of S-1
        file reads + H2D copies (ref remote.py:205-215 / local.py:208-212).  False -> caller takes the file path."""
        import torch.distributed as _dist
        transport = self.cache.get('transport', Transport.FILE)
        src_site = self.input.get('pretrained_site')
        if transport not in (Transport.NVLINK, Transport.NCCL) or src_site is None:
            return False
        if not (_dist.is_available() and _dist.is_initialized() and _dist.get_world_size() > 1):
            return False
        learner = self._get_learner_cls(rt.learner_cls)(trainer=rt.trainer, mp_pool=rt.mp_pool)
        arena = getattr(learner, 'arena', None)
        if arena is None or not hasattr(arena, 'broadcast_from'):
            return False
        mine = self.state['clientId'] == src_site
        if mine:
            rt.trainer.load_checkpoint(file_path=path)
        arena.broadcast_from(is_source=mine, model=learner.model)
        self.out['weights_broadcast'] = self.cache['_weights_broadcast'] = 'device'
        return True

    _ENTRY = {Phase.INIT_RUNS: '_enter_init_runs', Phase.NEXT_RUN: '_enter_next_run',
              Phase.PRE_COMPUTATION: '_enter_pre_computation'}
    _EXIT = {Phase.COMPUTATION: '_computation_round', Phase.SUCCESS: '_collect_results'}

    def compute(self, mp_pool, trainer_cls, dataset_cls=None, datahandle_cls=_DataHandle,
                learner_cls=_dSGDLearner, **kw):
        rt = _Round(mp_pool=mp_pool, trainer_cls=trainer_cls, dataset_cls=dataset_cls, datahandle_cls=datahandle_cls,
                    learner_cls=learner_cls)
        rt.trainer = trainer_cls(data_handle=datahandle_cls(
            cache=self.cache, input=self.input, state=self.state, dataloader_args=self._dataloader_args))
        self.out['phase'] = self.input.get('phase', Phase.INIT_RUNS)
        self._dispatch(self._ENTRY, self.out['phase'], rt)
        rt.learner = self._get_learner_cls(learner_cls)(trainer=rt.trainer, mp_pool=mp_pool)
        rt.modes = list(rt.learner.global_modes.values())
        self.out['mode'] = rt.learner.global_modes.get(self.state['clientId'], self.cache['mode'])
        self._dispatch(self._EXIT, self.out['phase'], rt)

    def _dispatch(self, table, key, rt):
        name = table.get(key)
        if name is not None:
            getattr(self, name)(rt)

    # ---- computation-phase rules: (guard, action), evaluated in order, all that match fire ----
    def _do_save_best(self, rt):
        self._sync_optimizer_state()
        rt.trainer.save_checkpoint(file_path=self.cache['log_dir'] + _sep + self.cache['best_nn_state'])

    def _do_save_resume_point(self, rt):
        """Phase 2 of a resume point: model + optimizer (+ the loader epoch) of this site at the end of ``epoch``.  Files of
        points older than the last COMMITTED one are pruned; the committed one stays until its successor is committed."""
        epoch = int(self.input['save_resume_point'])
        self._sync_optimizer_state()
        rt.trainer.save_checkpoint(file_path=self._resume_path(epoch))
        with open(self._resume_path(epoch, 'json'), 'w') as fp:
            _json.dump({'epoch': epoch, 'local_epoch': int(self.cache.get('local_epoch', 0))}, fp)
        engine_state = {k: self.cache[k] for k in ('powerSGD_state',) if k in self.cache}
        if engine_state:                                         # PowerSGD: error feedback, warm-start factors, iteration count
            import torch as _torch
            _torch.save(engine_state, self._resume_path(epoch, 'engine.pt'))
        self.out['resume_point_saved'] = epoch
        keep_from = self.cache.get('resume_committed')
        prefix = f"resume.{self.cache['task_id']}-{self.cache['split_ix']}.e"
        for name in _os.listdir(self.cache['log_dir']):
            if name.startswith(prefix):
                try:
                    e = int(name[len(prefix):].split('.')[0])
                except ValueError:
                    continue
                if keep_from is not None and e < int(keep_from):
                    _os.remove(_os.path.join(self.cache['log_dir'], name))

    def _note_resume_commit(self, rt):
        self.cache['resume_committed'] = int(self.input['resume_point_committed'])

    def _do_update(self, rt):
        self.out.update(**rt.learner.step())

    def _do_train(self, rt):
        # Lagging sites re-shuffle and keep contributing until *everyone* is waiting.
        it, out = rt.learner.to_reduce()
        self.out.update(**out)
        if it.get('averages') and it.get('metrics'):
            self.cache[Key.TRAIN_SERIALIZABLE].append(
                {'averages': it['averages'].serialize(), 'metrics': it['metrics'].serialize()})
            self.out.update(**rt.trainer.on_iteration_end(0, 0, it))

    def _do_validate(self, rt):
        self.out.update(**rt.trainer.validation_distributed(rt.dataset_cls))
        self.out[Key.TRAIN_SERIALIZABLE] = self.cache[Key.TRAIN_SERIALIZABLE]
        self.cache[Key.TRAIN_SERIALIZABLE] = []
        self.out['mode'] = Mode.TRAIN_WAITING

    def _do_test(self, rt):
        self.out.update(**rt.trainer.test_distributed(rt.dataset_cls))
        self.out['mode'] = self.cache['frozen_args']['mode']
        self.out['phase'] = Phase.NEXT_RUN_WAITING
        self._sync_optimizer_state()
        rt.trainer.save_checkpoint(file_path=self.cache['log_dir'] + _sep + self.cache['latest_nn_state'])
        _utils.save_cache(self.cache, self.cache['log_dir'])

    _COMPUTATION_RULES = (
        (lambda self, rt: bool(self.input.get('save_current_as_best')), '_do_save_best'),
        (lambda self, rt: self.input.get('resume_point_committed') is not None, '_note_resume_commit'),
        (lambda self, rt: self.input.get('save_resume_point') is not None, '_do_save_resume_point'),
        (lambda self, rt: bool(self.input.get('update')), '_do_update'),
        (lambda self, rt: any(m == Mode.TRAIN for m in rt.modes), '_do_train'),
        (lambda self, rt: bool(rt.modes) and all(m == Mode.VALIDATION for m in rt.modes), '_do_validate'),
        (lambda self, rt: bool(rt.modes) and all(m == Mode.TEST for m in rt.modes), '_do_test'),
    )

    def _computation_round(self, rt):
        for guard, action in self._COMPUTATION_RULES:
            if guard(self, rt):
                getattr(self, action)(rt)

    def _sync_optimizer_state(self):
        """Sharded (two-shot / NVLS) optimizer moments are collected before a global checkpoint;
        every site reaches these save points in the same round, so the collective is matched."""
        arena = self.cache.get('_arena')
        if arena is not None:
            arena.gather_state()

    def _collect_results(self, rt=None):
        """Final round: pick up the results zip broadcast by the aggregator (retry x3)."""
        name = f"{self.input['results_zip']}.zip"
        src = f"{self.state['baseDirectory']}{_sep}{name}"
        for attempt in range(3):
            _time.sleep(attempt * float(self.cache.get('zip_retry_seconds', 1.0)))
            if _os.path.exists(src):
                _shutil.copy(src, f"{self.state['outputDirectory']}{_sep}{name}")
                break

    def _get_learner_cls(self, learner_cls):
        """``agg_engine`` x ``transport`` -> learner class; unknown engines use ``learner_cls``."""
        engine = self.cache.get('agg_engine')
        transport = self.cache.get('transport', Transport.FILE)
        if transport in (Transport.NVLINK, Transport.NCCL):
            from ...parallel import nvlink_learner as _nv
            table = {AGG_Engine.dSGD: _nv.NvlinkLearner, AGG_Engine.powerSGD: _nv.NvlinkPowerSGDLearner,
                     AGG_Engine.rankDAD: _nv.NvlinkDADLearner}
            if engine in table:
                return table[engine]
        return _engine_learners().get(engine, learner_cls)

    def __call__(self, *args, **kwargs):
        try:
            self.compute(*args, **kwargs)
            return {'output': self.out}
        except Exception:
            _tback.print_exc()
            raise Exception(self.out)
