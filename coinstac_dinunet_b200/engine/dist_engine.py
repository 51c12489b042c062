"""One-process-per-site round engine over ``torch.distributed`` (rank r == site ``local<r>``;
rank 0 additionally hosts the aggregator).

This is the B200 deployment shape: launched with ``torch.distributed.run`` (NCCL on GPUs, gloo on
CPU), every rank owns one GPU and its node cache.  The JSON control plane travels through
``gather_object`` / ``broadcast_object_list`` - tiny and, with the NVLink transport, touched once
per epoch; the tensor data plane never goes through here (it is inside the fused kernel).  Files
that still exist in the protocol (pre-trained weights broadcast, results zip, and the ``*.npy``
payloads of the file transport) are shipped by rank 0 on the node-local filesystem exactly like
the in-process emulator does.
"""
import os as _os
import time as _time

import torch as _torch
import torch.distributed as _dist

from .emulator import _clear_files, _copy_tree_flat, node_state, unwrap_spec


AFFINITY = {}      # report of utils.affinity.pin_to_gpu for this process (empty on CPU)


def init_process_group(backend=None):
    """Idempotent ``init_process_group`` from the torchrun environment (``127.0.0.1`` default)."""
    if _dist.is_initialized():
        return
    _os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    _os.environ.setdefault('MASTER_PORT', '29512')
    _os.environ.setdefault('RANK', '0')
    _os.environ.setdefault('WORLD_SIZE', '1')
    cuda = _torch.cuda.is_available()
    backend = backend or ('nccl' if cuda else 'gloo')
    kw = {}
    if backend == 'nccl':
        local = int(_os.environ.get('LOCAL_RANK', '0'))
        _torch.cuda.set_device(local)
        kw['device_id'] = _torch.device('cuda', local)
        from ..utils.affinity import pin_to_gpu           # before any pinned allocation: NUMA-local staging buffers
        world_local = int(_os.environ.get('LOCAL_WORLD_SIZE', _os.environ.get('WORLD_SIZE', '1')))
        AFFINITY.update(pin_to_gpu(local, ranks_per_node=max(1, (world_local + 1) // 2)))
    _dist.init_process_group(backend=backend, **kw)


def _jsonish(obj):
    """Make an output dict picklable/JSON-like (enum keys -> str, FrozenDict -> dict)."""
    if isinstance(obj, dict):
        return {str(k): _jsonish(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return [_jsonish(v) for v in obj]
    if isinstance(obj, str):
        return str(obj)
    return obj


class DistEngine:
    def __init__(self, work_dir, inputspec=None, clear_transfer=True, group=None):
        init_process_group()
        self.group = group
        self.rank, self.world = _dist.get_rank(group), _dist.get_world_size(group)
        # The JSON control plane rides a gloo group when the data plane is NCCL: object collectives over NCCL stage
        # their pickles through device tensors and synchronise the stream - a millisecond per round that the (tiny)
        # messages do not need.
        self.ctl = group
        try:
            if _dist.get_backend(group) == 'nccl' and self.world > 1 and _os.environ.get('COINN_CTL_GLOO', '1') == '1':
                ranks = _dist.get_process_group_ranks(group) if group is not None else None
                self.ctl = _dist.new_group(ranks=ranks, backend='gloo')
        except Exception:
            self.ctl = group
        self._mailbox = self._open_mailbox()
        self.work_dir = str(work_dir)
        self.site_ids = [f'local{i}' for i in range(self.world)]
        self.site = self.site_ids[self.rank]
        self.clear_transfer = clear_transfer
        self.state = node_state(self.work_dir, self.site)
        self.cache = {}
        spec = inputspec[self.rank] if isinstance(inputspec, (list, tuple)) else (inputspec or {})
        self.input = unwrap_spec(spec)
        self.remote_state = node_state(self.work_dir, 'remote') if self.rank == 0 else None
        self.remote_cache = {} if self.rank == 0 else None
        self.all_site_states = [node_state(self.work_dir, s) for s in self.site_ids] if self.rank == 0 else None
        self.trace, self.round, self.timings = [], 0, []

    def _open_mailbox(self):
        """Single-node runs exchange the per-round JSON through shared memory (``engine/shm_plane.py``); multi-node runs,
        or ``COINN_CTL_SHM=0``, keep the ``torch.distributed`` object collectives."""
        import atexit
        if self.world == 1 or _os.environ.get('COINN_CTL_SHM', '1') != '1':
            return None
        local_world = int(_os.environ.get('LOCAL_WORLD_SIZE', self.world))
        try:
            hosts = [None] * self.world
            _dist.all_gather_object(hosts, _os.uname().nodename, group=self.ctl)
            if local_world != self.world or len(set(hosts)) != 1:
                return None
            from .shm_plane import ShmMailbox
            box, mb, ok = [None], None, True
            if self.rank == 0:
                try:
                    mb = ShmMailbox(None, 0, self.world, create=True,
                                    slot_bytes=int(_os.environ.get('COINN_CTL_SLOT_BYTES', 1 << 20)))
                    box[0] = mb.name
                except Exception:
                    ok = False
            _dist.broadcast_object_list(box, src=0, group=self.ctl)
            if self.rank != 0 and box[0] is not None:
                try:
                    mb = ShmMailbox(box[0], self.rank, self.world, create=False,
                                    slot_bytes=int(_os.environ.get('COINN_CTL_SLOT_BYTES', 1 << 20)))
                except Exception:
                    ok = False
            votes = [None] * self.world                     # every rank must have the segment mapped, or nobody uses it
            _dist.all_gather_object(votes, bool(ok and mb is not None), group=self.ctl)
            if not all(votes):
                if mb is not None:
                    mb.close()
                return None
            mb.verify_pids()                                # every rank has attached (the vote above was the rendezvous)
            atexit.register(mb.close)
            grp = self.ctl

            def send_big(obj, peer):
                _dist.send_object_list([obj], dst=peer, group=grp)

            def recv_big(peer):
                box = [None]
                _dist.recv_object_list(box, src=peer, group=grp)
                return box[0]

            def bcast_big(obj):
                box = [obj]
                _dist.broadcast_object_list(box, src=0, group=grp)
                return box[0]
            mb.send_big, mb.recv_big, mb.bcast_big = send_big, recv_big, bcast_big
            return mb
        except Exception:
            return None

    def _gather(self, obj):
        if self.world == 1:                      # a collective of one still pickles, stages and synchronises: ~0.5 ms per round
            return [obj]
        if self._mailbox is not None:
            return self._mailbox.gather(obj)
        gathered = [None] * self.world if self.rank == 0 else None
        _dist.gather_object(obj, gathered, dst=0, group=self.ctl)
        return gathered

    def _broadcast(self, payload):
        if self.world == 1:
            return payload
        if self._mailbox is not None:
            return self._mailbox.broadcast(payload)
        box = [payload]
        _dist.broadcast_object_list(box, src=0, group=self.ctl)
        return box[0]

    def step(self, local_fn, remote_fn):
        try:
            return self._step(local_fn, remote_fn)
        except BaseException:
            if self._mailbox is not None:                                   # wake the ranks that wait for this one
                self._mailbox.abort()
            raise

    def _step(self, local_fn, remote_fn):
        t0 = _time.time()
        out = local_fn(self.site, self.cache, self.input, self.state)['output']
        gathered = self._gather(_jsonish(out))                              # doubles as file barrier
        payload = None
        if self.rank == 0:
            for st in self.all_site_states:
                src = st['transferDirectory']
                _copy_tree_flat(src, _os.path.join(self.remote_state['baseDirectory'], st['clientId']))
                if self.clear_transfer:
                    _clear_files(src)
            res = remote_fn(self.remote_cache, dict(zip(self.site_ids, gathered)), self.remote_state)
            for st in self.all_site_states:
                _copy_tree_flat(self.remote_state['transferDirectory'], st['baseDirectory'])
            if self.clear_transfer:
                _clear_files(self.remote_state['transferDirectory'])
            payload = (_jsonish(res['output']), bool(res.get('success')))
        remote_out, success = self._broadcast(payload)
        self.input = dict(remote_out)
        self.trace.append({'site': (str(out.get('phase')), str(out.get('mode'))),
                           'remote': str(remote_out.get('phase'))})
        self.timings.append(_time.time() - t0)
        self.round += 1
        return success

    def run(self, local_fn, remote_fn, max_rounds=100000):
        while self.round < max_rounds:
            if self.step(local_fn, remote_fn):
                return self.round
        raise RuntimeError(f'engine did not converge within {max_rounds} rounds')

    def run_nodes(self, trainer_cls, dataset_cls=None, datahandle_cls=None, local_kw=None, remote_kw=None,
                  mp_pool=None, max_rounds=100000):
        from ..data import COINNDataHandle
        from ..distrib.nodes import COINNLocal, COINNRemote
        local_kw, remote_kw = dict(local_kw or {}), dict(remote_kw or {})
        dh = datahandle_cls or COINNDataHandle

        def local_fn(site, cache, inp, state):
            return COINNLocal(cache=cache, input=inp, state=state, **local_kw)(mp_pool, trainer_cls, dataset_cls, dh)

        def remote_fn(cache, inp, state):
            return COINNRemote(cache=cache, input=inp, state=state, **remote_kw)(mp_pool, trainer_cls)

        return self.run(local_fn, remote_fn, max_rounds=max_rounds)
