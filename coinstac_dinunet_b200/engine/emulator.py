"""In-process COINSTAC-compatible round engine (the reference relies on an *external* engine,
SURVEY §2.6; this is the stand-in that makes the protocol testable and runnable anywhere).

One *round* = every site computes once -> engine ships each site's ``transferDirectory`` to
``<remote baseDirectory>/<site>/`` and its JSON output to the aggregator -> aggregator computes ->
engine ships the aggregator's ``transferDirectory`` into every site's ``baseDirectory`` and its
JSON output back.  Node caches are plain dicts owned by the engine (they hold live modules,
iterators, arenas - SURVEY fact 4).  Directory layout mirrors the COINSTAC simulator::

    <work>/input/<node>/simulatorRun      baseDirectory
    <work>/output/<node>/simulatorRun     outputDirectory
    <work>/transfer/<node>/simulatorRun   transferDirectory
"""
import os as _os
import shutil as _shutil
import time as _time


def _copy_tree_flat(src, dst):
    """Copy every regular file of ``src`` into ``dst`` (created on demand).  An empty ``src`` - every round of the device
    transports - costs one ``scandir`` and nothing else."""
    try:
        with _os.scandir(src) as it:
            files = [e.name for e in it if e.is_file()]
    except (FileNotFoundError, NotADirectoryError):
        return 0
    if not files:
        return 0
    _os.makedirs(dst, exist_ok=True)
    for name in files:
        _shutil.copy(_os.path.join(src, name), _os.path.join(dst, name))
    return len(files)


def _clear_files(folder):
    try:
        with _os.scandir(folder) as it:
            files = [e.path for e in it if e.is_file()]
    except (FileNotFoundError, NotADirectoryError):
        return
    for path in files:
        _os.remove(path)


def node_state(work_dir, node_id):
    """The ``state`` dict a COINSTAC engine hands to a node (ref site_runner.py:17-24)."""
    st = {'clientId': node_id}
    for key, top in (('baseDirectory', 'input'), ('outputDirectory', 'output'),
                     ('transferDirectory', 'transfer')):
        st[key] = _os.path.join(work_dir, top, node_id, 'simulatorRun')
        _os.makedirs(st[key], exist_ok=True)
    return st


def unwrap_spec(spec):
    """``{'k': {'value': v}}`` (inputspec.json style) or plain ``{'k': v}`` -> ``{'k': v}``."""
    return {k: (v['value'] if isinstance(v, dict) and set(v) == {'value'} else v)
            for k, v in (spec or {}).items()}


class InProcessEngine:
    """Drive ``n_sites`` local nodes and one remote node through the round protocol.

    ``local_fn(site_id, cache, input, state) -> {'output': dict}``
    ``remote_fn(cache, input, state) -> {'output': dict, 'success': bool}``
    ``inputspec`` is either one dict shared by all sites or a list with one dict per site.
    """

    def __init__(self, work_dir, n_sites=2, inputspec=None, site_ids=None, clear_transfer=True):
        self.work_dir = str(work_dir)
        self.site_ids = list(site_ids) if site_ids else [f'local{i}' for i in range(n_sites)]
        self.remote_id = 'remote'
        self.clear_transfer = clear_transfer
        self.site_state = {s: node_state(self.work_dir, s) for s in self.site_ids}
        self.remote_state = node_state(self.work_dir, self.remote_id)
        self.site_cache = {s: {} for s in self.site_ids}
        self.remote_cache = {}
        specs = inputspec if isinstance(inputspec, (list, tuple)) else [inputspec or {}] * len(self.site_ids)
        self.site_input = {s: unwrap_spec(spec) for s, spec in zip(self.site_ids, specs)}
        self.trace = []          # per round: sites' (phase, mode), remote phase, global modes
        self.round = 0
        self.timings = []
        self.fault_hook = None   # callable(round, site_id) -> raise to inject a failure
        self._partial = {}       # outputs of sites that already finished the round in flight (retry-safe)

    # ------------------------------------------------------------------ shipping
    def _ship_site_to_remote(self, site):
        src = self.site_state[site]['transferDirectory']
        _copy_tree_flat(src, _os.path.join(self.remote_state['baseDirectory'], site))
        if self.clear_transfer:
            _clear_files(src)

    def _ship_remote_to_sites(self):
        src = self.remote_state['transferDirectory']
        for site in self.site_ids:
            _copy_tree_flat(src, self.site_state[site]['baseDirectory'])
        if self.clear_transfer:
            _clear_files(src)

    # --------------------------------------------------------------------- rounds
    def step(self, local_fn, remote_fn):
        """Run one full round; returns ``success`` reported by the aggregator."""
        t0 = _time.time()
        # A node call is not idempotent (it consumes a batch / applies an update), so when a round is retried
        # after a failure the sites that had already completed it are NOT called again: their outputs (and
        # shipped files) are kept in `_partial` until the round commits.
        site_out = self._partial
        for site in self.site_ids:
            if site in site_out:
                continue
            if self.fault_hook is not None:
                self.fault_hook(self.round, site)
            res = local_fn(site, self.site_cache[site], self.site_input[site], self.site_state[site])
            site_out[site] = res['output']
            self._ship_site_to_remote(site)
        res = remote_fn(self.remote_cache, site_out, self.remote_state)
        self._partial = {}
        self._ship_remote_to_sites()
        remote_out = res['output']
        for site in self.site_ids:
            self.site_input[site] = dict(remote_out)
        self.trace.append({
            'sites': {s: (str(o.get('phase')), str(o.get('mode'))) for s, o in site_out.items()},
            'remote': str(remote_out.get('phase')),
            'modes': {k: str(v) for k, v in (remote_out.get('global_modes') or {}).items()},
        })
        self.timings.append(_time.time() - t0)
        self.round += 1
        return bool(res.get('success'))

    def run(self, local_fn, remote_fn, max_rounds=100000):
        """Rounds until the aggregator reports success.  Returns the number of rounds."""
        while self.round < max_rounds:
            if self.step(local_fn, remote_fn):
                return self.round
        raise RuntimeError(f'engine did not converge within {max_rounds} rounds')

    # ---------------------------------------------------------------- convenience
    def run_nodes(self, trainer_cls, dataset_cls=None, datahandle_cls=None, local_kw=None, remote_kw=None,
                  mp_pool=None, max_rounds=100000, learner_cls=None, reducer_cls=None):
        """Wire ``COINNLocal`` / ``COINNRemote`` into the engine and run to completion."""
        from ..data import COINNDataHandle
        from ..distrib.nodes import COINNLocal, COINNRemote
        local_kw, remote_kw = dict(local_kw or {}), dict(remote_kw or {})
        dh = datahandle_cls or COINNDataHandle

        def local_fn(site, cache, inp, state):
            node = COINNLocal(cache=cache, input=inp, state=state, **local_kw)
            extra = {'learner_cls': learner_cls} if learner_cls else {}
            return node(mp_pool, trainer_cls, dataset_cls, dh, **extra)

        def remote_fn(cache, inp, state):
            node = COINNRemote(cache=cache, input=inp, state=state, **remote_kw)
            extra = {'reducer_cls': reducer_cls} if reducer_cls else {}
            return node(mp_pool, trainer_cls, **extra)

        return self.run(local_fn, remote_fn, max_rounds=max_rounds)
