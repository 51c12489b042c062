"""Flat parameter / gradient / optimizer-state arenas and the fused dSGD step.

``DistArena`` re-homes the parameters of a model into one flat fp32 buffer (``p.data`` become
views), pre-binds ``p.grad`` to views of one flat *symmetric* gradient buffer that autograd then
accumulates into in place (this removes the reference's per-parameter D2H/H2D copies,
tensorutils.py:44-47 / learner.py:25-26), and keeps Adam/SGD state in flat buffers.  One call
``reduce_and_step()`` launches ``fused_reduce_opt.cu`` which averages the gradients of all sites
over NVLink, applies the optimizer and re-zeroes the gradient buffer - the whole of the reference's
``to_reduce -> COINNReducer.reduce -> step`` round trip.

Variant choice per bucket (``variant='auto'``): one-shot below ``config.ONE_SHOT_MAX_BYTES``,
otherwise NVLS when the allocation has a multicast alias, else two-shot.
Backends: ``'nvlink'`` fused kernel (product) | ``'nccl'`` all-reduce + fused local step (the
baseline the spec names) | ``'torch'`` all-reduce + torch optimizer (CPU/gloo, used by the tests).
"""
import ctypes as _C

import torch as _torch
import torch.distributed as _dist

from .. import config as _conf
from .symm import SymmetricBuffer, _rank, _world

_ALIGN = 8          # elements; keeps every parameter 16-byte aligned in fp32 and in a bf16 shadow
_VARIANTS = {'one_shot': 0, 'two_shot': 1, 'nvls': 2}
_OPT_KINDS = {'adam': 0, 'adamw': 1, 'sgd': 2, 'none': 3}
_WIRE = {'f32': (0, None), 'fp32': (0, None), 'float32': (0, None), None: (0, None),
         'bf16': (1, _torch.bfloat16), 'bfloat16': (1, _torch.bfloat16),
         'f16': (2, _torch.float16), 'fp16': (2, _torch.float16), 'float16': (2, _torch.float16)}
_DEFAULT_TIMEOUT_MS = 30000


def wire_dtype_for(cache):
    """``cache['grad_dtype']`` ('bf16' / 'fp16') or the reference's ``precision_bits`` (16 -> float16, learner.py:17)."""
    if cache.get('grad_dtype'):
        return str(cache['grad_dtype'])
    return 'f16' if int(cache.get('precision_bits', 32) or 32) == 16 else 'f32'


def _round_up(n, a):
    return (n + a - 1) // a * a


def describe_optimizer(opt):
    """Hyper-parameters of a torch optimizer that the fused kernel can reproduce, or None."""
    if len(opt.param_groups) != 1:
        return None
    g = opt.param_groups[0]
    if isinstance(opt, _torch.optim.AdamW) or isinstance(opt, _torch.optim.Adam):
        if g.get('amsgrad') or g.get('maximize'):
            return None
        kind = 'adamw' if (isinstance(opt, _torch.optim.AdamW) or g.get('decoupled_weight_decay')) else 'adam'
        return dict(kind=kind, lr=float(g['lr']), beta1=float(g['betas'][0]), beta2=float(g['betas'][1]),
                    eps=float(g['eps']), weight_decay=float(g['weight_decay']), momentum=0.0, nesterov=0)
    if isinstance(opt, _torch.optim.SGD):
        if g.get('maximize') or g.get('dampening', 0) != 0:
            return None
        return dict(kind='sgd', lr=float(g['lr']), beta1=0.0, beta2=0.0, eps=0.0,
                    weight_decay=float(g['weight_decay']), momentum=float(g['momentum']),
                    nesterov=int(bool(g['nesterov'])))
    return None


class DistArena:
    def __init__(self, model, optimizer, device=None, group=None, backend='auto', shadow_bf16=False,
                 variant='auto', bucket_bytes=None, grad_dtype='f32', timeout_ms=None):
        self.model, self.optimizer, self.group = model, optimizer, group
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.device = _torch.device(device) if device is not None else self.params[0].device
        self.world, self.rank = _world(group), _rank(group)
        self.variant = variant
        self.hyper = describe_optimizer(optimizer)

        if backend == 'auto':
            backend = 'nvlink' if self.device.type == 'cuda' else 'torch'
        if backend in ('nvlink', 'nccl') and (self.device.type != 'cuda' or self.hyper is None):
            backend = 'torch'
        self.backend = backend

        # ---- layout -------------------------------------------------------------------------
        self.offsets, off = [], 0
        for p in self.params:
            self.offsets.append(off)
            off += _round_up(p.numel(), _ALIGN)
        self.numel = _round_up(max(off, _ALIGN), 4 * max(self.world, 1))
        self.bucket_bytes = bucket_bytes

        sym = backend == 'nvlink'
        mk = (lambda n, dt: SymmetricBuffer(n, dt, self.device, group)) if sym else \
            (lambda n, dt: _Local(n, dt, self.device))
        # gradient exchange dtype on the NVLink wire: fp32 peers read the arena itself; 16-bit (precision_bits=16)
        # packs the fp32 arena into a symmetric half-size wire buffer inside the fused kernel (half the link bytes)
        self.grad_code, wire_dt = _WIRE[grad_dtype if grad_dtype in _WIRE else str(grad_dtype).lower()]
        if not (sym and self.world > 1):
            self.grad_code, wire_dt = 0, None
        self.grad_buf = mk(self.numel, _torch.float32) if self.grad_code == 0 else _Local(self.numel, _torch.float32, self.device)
        self.wire_buf = mk(self.numel, wire_dt) if self.grad_code else None
        self.param_buf = mk(self.numel, _torch.float32)
        self.shadow_buf = mk(self.numel, _torch.bfloat16) if shadow_bf16 else None
        self.flat_grad, self.flat_param = self.grad_buf.local, self.param_buf.local
        self.m = _torch.zeros(self.numel, dtype=_torch.float32, device=self.device)
        self.v = _torch.zeros(self.numel, dtype=_torch.float32, device=self.device)
        self.step_count = _torch.zeros(1, dtype=_torch.int32, device=self.device)
        # learning rate as a device scalar: a captured CUDA graph reads it at replay time, so LR schedulers keep
        # working under ``cuda_graph=True`` (``sync_lr`` refreshes it when ``param_groups[0]['lr']`` changed)
        self.lr_dev = _torch.zeros(1, dtype=_torch.float32, device=self.device)
        self._lr_uploaded = None
        self.steps_done = 0     # optimizer steps taken through this arena (host mirror of step_count)
        self.host_step = 0      # Adam's `step` as torch.optim would report it

        if backend in ('nvlink', 'nccl'):
            from ..ops import native as _nat
            self._nat = _nat
            slots = _nat.lib().coinn_fused_flag_slots()
            self.flags = mk(slots, _torch.int32)
            self.epoch = _torch.zeros(_nat.lib().coinn_fused_max_blocks(), dtype=_torch.int32, device=self.device)
            self.ticket = _torch.zeros(1, dtype=_torch.int32, device=self.device)
            self.error = _torch.zeros(1, dtype=_torch.int32, device=self.device)    # watchdog word of the barrier
            import os as _os
            self.timeout_ms = int(timeout_ms if timeout_ms is not None else
                                  float(_os.environ.get('COINN_BARRIER_TIMEOUT_S', _DEFAULT_TIMEOUT_MS / 1e3)) * 1e3)
            self._args_cache = {}

        self._bind()
        self.import_optimizer_state()

    # ------------------------------------------------------------------------------ binding
    def _bind(self):
        """Move parameter storage into the arena and pre-bind ``.grad`` views."""
        with _torch.no_grad():
            for p, off in zip(self.params, self.offsets):
                n = p.numel()
                view = self.flat_param[off:off + n].view(p.shape)
                view.copy_(p.data.to(self.device, _torch.float32))
                p.data = view
                p.grad = self.flat_grad[off:off + n].view(p.shape)
            if self.shadow_buf is not None:
                self.shadow_buf.local.copy_(self.flat_param)
        if self.world > 1 and self.param_buf.__class__ is SymmetricBuffer:
            _torch.cuda.synchronize(self.device)
            self.param_buf.barrier()

    def rebind_grads(self):
        """Re-attach ``p.grad`` to the arena (after ``optimizer.zero_grad(set_to_none=True)``)."""
        for p, off in zip(self.params, self.offsets):
            if p.grad is None or p.grad.data_ptr() != self.flat_grad.data_ptr() + off * 4:
                p.grad = self.flat_grad[off:off + p.numel()].view(p.shape)

    def refresh_from_params(self):
        """After a checkpoint load wrote new values through ``p.data``: nothing to copy (the views
        ARE the arena) unless a loader replaced the tensors; then re-bind."""
        with _torch.no_grad():
            for p, off in zip(self.params, self.offsets):
                if p.data_ptr() != self.flat_param.data_ptr() + off * 4:
                    self.flat_param[off:off + p.numel()].view(p.shape).copy_(p.data)
                    p.data = self.flat_param[off:off + p.numel()].view(p.shape)
            if self.shadow_buf is not None:
                self.shadow_buf.local.copy_(self.flat_param)
        self.rebind_grads()
        self.import_optimizer_state()

    # ------------------------------------------------------------------- optimizer state bridge
    def _publish_state_views(self):
        """Expose the arena through ``optimizer.state`` so ``state_dict()`` stays torch-compatible."""
        if self.hyper is None or self.backend == 'torch':
            return
        for p, off in zip(self.params, self.offsets):
            n = p.numel()
            st = self.optimizer.state[p]
            if self.hyper['kind'] == 'sgd':
                st['momentum_buffer'] = self.m[off:off + n].view(p.shape)
            else:
                st['step'] = _torch.tensor(float(self.host_step))
                st['exp_avg'] = self.m[off:off + n].view(p.shape)
                st['exp_avg_sq'] = self.v[off:off + n].view(p.shape)

    def import_optimizer_state(self):
        """Copy existing torch optimizer state (e.g. from a loaded checkpoint) into the arena."""
        if self.hyper is None or self.backend == 'torch':
            return
        step = 0
        with _torch.no_grad():
            for p, off in zip(self.params, self.offsets):
                st = self.optimizer.state.get(p, {})
                n = p.numel()
                for key, buf in (('exp_avg', self.m), ('exp_avg_sq', self.v), ('momentum_buffer', self.m)):
                    t = st.get(key)
                    if t is not None and t.data_ptr() != buf.data_ptr() + off * 4:
                        buf[off:off + n].view(p.shape).copy_(t.to(self.device, _torch.float32))
                if 'step' in st:
                    step = max(step, int(float(st['step'])))
        if step:
            self.step_count.fill_(step)
            self.host_step = step
        self._publish_state_views()

    def _launch_units(self):
        """(offset, numel) of every fused launch of one step: the buckets when overlap is on, else the whole arena."""
        ov = getattr(self, '_overlap', None)
        if ov:
            return [(b['offset'], b['numel']) for b in ov['buckets']]
        return [(0, self.numel)]

    def owner_ranges(self):
        """[(rank, lo, hi)] element ranges whose optimizer moments live on exactly one rank: two-shot / NVLS launches
        shard every launch unit as ``ceil(nvec / S)`` 4-element vectors per rank (fused_reduce_opt.cu), one-shot units
        are replicated and do not appear."""
        out = []
        if self.world == 1 or self.backend != 'nvlink':
            return out
        for off, numel in self._launch_units():
            if self._pick_variant(numel * 4) == 'one_shot':
                continue
            nvec = numel // 4
            shard = -(-nvec // self.world)
            for q in range(self.world):
                lo, hi = min(q * shard, nvec), min((q + 1) * shard, nvec)
                if hi > lo:
                    out.append((q, off + 4 * lo, off + 4 * hi))
        return out

    def gather_state(self):
        """Two-shot/NVLS keep optimizer moments sharded (rank r owns shard r of every launch unit).  Before a
        checkpoint - or before switching to full-range local updates - every rank collects the other shards."""
        if self.world > 1 and self.backend == 'nvlink':
            for q, lo, hi in self.owner_ranges():
                src = _dist.get_global_rank(self.group, q) if self.group is not None else q
                for buf in (self.m, self.v):
                    _dist.broadcast(buf[lo:hi], src=src, group=self.group)
        self._publish_state_views()

    # --------------------------------------------------------------- C5: device-to-device broadcast
    def broadcast_from(self, is_source, model=None):
        """Make every site a copy of the one site with ``is_source=True``: fp32 master parameters, optimizer moments,
        step counter and (optionally) the float buffers of ``model`` (BatchNorm running statistics).  On the NVLink
        backend the bulk moves as peer-to-peer copies out of the source's symmetric buffers (parameter arena directly;
        moments / buffers staged through the gradient arena, which is idle and zero between steps) - no file, no host
        bounce, no NCCL on the payload.  Other backends use ``torch.distributed.broadcast``."""
        if self.world == 1:
            return 0
        dev = self.device if self.device.type == 'cuda' else _torch.device('cpu')
        who = _torch.tensor([self.rank if is_source else -1], dtype=_torch.int64, device=dev)
        _dist.all_reduce(who, op=_dist.ReduceOp.MAX, group=self.group)
        src = int(who.item())
        if src < 0:
            raise RuntimeError('broadcast_from: no site declared itself the source')
        gsrc = _dist.get_global_rank(self.group, src) if self.group is not None else src
        bufs = [b for b in (model.buffers() if model is not None else []) if b.is_floating_point()]
        ints = [b for b in (model.buffers() if model is not None else []) if not b.is_floating_point()]
        symmetric = self.backend == 'nvlink' and isinstance(self.param_buf, SymmetricBuffer)
        if symmetric:
            # every site must be able to alias the source's buffers, or nobody takes the peer-copy route
            try:
                ok = self.param_buf.peer_tensor(src).numel() == self.numel
            except Exception:
                ok = False
            votes = _torch.tensor([int(ok)], dtype=_torch.int64, device=dev)
            _dist.all_reduce(votes, op=_dist.ReduceOp.MIN, group=self.group)
            symmetric = bool(int(votes.item()))
        staging = self.grad_buf if (symmetric and isinstance(self.grad_buf, SymmetricBuffer)) else None

        def sync():
            _torch.cuda.synchronize(self.device)
            self.param_buf.barrier()

        with _torch.no_grad():
            if symmetric:
                sync()                                           # the source has finished loading its checkpoint
                if self.rank != src:
                    self.flat_param.copy_(self.param_buf.peer_tensor(src))
                sync()
            else:
                _dist.broadcast(self.flat_param, src=gsrc, group=self.group)
            payload = [self.m, self.v] if self.backend != 'torch' else []
            if bufs:
                flat_b = _torch.cat([b.detach().reshape(-1).float() for b in bufs])
                payload.append(flat_b)
            for t in payload:
                n = t.numel()
                if staging is not None and n <= staging.numel:
                    if self.rank == src:
                        staging.local[:n].copy_(t)
                    sync()
                    if self.rank != src:
                        t.copy_(staging.peer_tensor(src)[:n])
                    sync()
                    if self.rank == src:
                        staging.local[:n].zero_()                # the gradient arena must be zero when a step starts
                else:
                    _dist.broadcast(t, src=gsrc, group=self.group)
            if bufs:
                off = 0
                for b in bufs:
                    b.copy_(flat_b[off:off + b.numel()].view_as(b).to(b.dtype))
                    off += b.numel()
            for b in ints:
                _dist.broadcast(b, src=gsrc, group=self.group)
            if self.backend != 'torch':
                _dist.broadcast(self.step_count, src=gsrc, group=self.group)
                self.host_step = int(self.step_count.item())
            if self.shadow_buf is not None:
                self.shadow_buf.local.copy_(self.flat_param)
            if symmetric:
                sync()
        self._publish_state_views()
        if self.backend == 'torch':          # CPU / exotic optimizers: ship the torch optimizer state as an object
            box = [self.optimizer.state_dict() if self.rank == src else None]
            _dist.broadcast_object_list(box, src=gsrc, group=self.group)
            if self.rank != src:
                self.optimizer.load_state_dict(box[0])
        return src

    # ------------------------------------------------------------------------------- the step
    def _pick_variant(self, nbytes):
        v = self.variant
        if v == 'auto':
            if self.world == 1 or nbytes <= _conf.ONE_SHOT_MAX_BYTES:
                v = 'one_shot'
            elif self.world > 2 and self._exchange_buf.multicast_ptr and self.param_buf.multicast_ptr:
                v = 'nvls'          # in-switch reduction pays from 4 sites up; at 2 sites two-shot is as fast or faster
            else:                   # (profiles/r2/reduce_sweep_n2.json: 60 vs 77 us at 16 MB, 2.0 vs 2.8 ms at 1 GB)
                v = 'two_shot'
        if v == 'nvls' and not (self._exchange_buf.multicast_ptr and self.param_buf.multicast_ptr):
            v = 'two_shot'
        return v

    @property
    def _exchange_buf(self):
        """The buffer peers read gradients from: the fp32 arena itself, or the 16-bit wire buffer."""
        return self.wire_buf if self.wire_buf is not None else self.grad_buf

    def check_health(self):
        """Raise if the barrier watchdog of the fused kernel fired (a site never arrived).  One 4-byte read: call it
        at round boundaries, not per step."""
        if self.backend != 'nvlink' or self.world == 1:
            return
        code = int(self.error.item())
        if code:
            raise RuntimeError(f'fused reduce: site (rank) {code - 1} did not reach the cross-GPU barrier within '
                               f'{self.timeout_ms} ms - treating it as failed (SURVEY §5.3)')

    def _fused_args(self, offset, numel, variant, world, grad_scale, zero_grads, bump):
        nat, h = self._nat, self.hyper
        key = (offset, numel, variant, world, zero_grads, bump)
        a = self._args_cache.get(key)
        if a is None:
            a = nat.FusedArgs()
            ex = self._exchange_buf
            for r in range(world):
                a.grad_ptrs[r] = ex.peer_ptrs[r] if world > 1 else self.flat_grad.data_ptr()
                a.param_ptrs[r] = self.param_buf.peer_ptrs[r] if world > 1 else self.flat_param.data_ptr()
                a.flag_ptrs[r] = self.flags.peer_ptrs[r] if world > 1 else self.flags.local.data_ptr()
                if self.shadow_buf is not None:
                    a.shadow_ptrs[r] = self.shadow_buf.peer_ptrs[r] if world > 1 else self.shadow_buf.local.data_ptr()
            if world > 1:
                a.grad_mc = ex.multicast_ptr or None
                a.param_mc = self.param_buf.multicast_ptr or None
                a.shadow_mc = (self.shadow_buf.multicast_ptr or None) if self.shadow_buf is not None else None
            a.m, a.v = self.m.data_ptr(), self.v.data_ptr()
            a.epoch, a.step, a.ticket = self.epoch.data_ptr(), self.step_count.data_ptr(), self.ticket.data_ptr()
            a.offset, a.numel = offset, numel
            a.rank, a.world = (self.rank if world > 1 else 0), world
            a.variant, a.opt_kind = _VARIANTS[variant], _OPT_KINDS[h['kind']]
            a.grad_dtype = self.grad_code if world > 1 else 0
            a.grad32 = self.flat_grad.data_ptr()
            a.error, a.timeout_ms = self.error.data_ptr(), self.timeout_ms
            a.zero_grads, a.bump_step, a.nesterov = int(zero_grads), int(bump), h['nesterov']
            self._args_cache[key] = a
        g = self.optimizer.param_groups[0]   # live values: lr schedulers keep working (eager: by value; graphs: lr_dev)
        a.lr = float(g['lr'])
        a.lr_ptr = self.lr_dev.data_ptr()
        self.sync_lr()
        a.beta1, a.beta2, a.eps = h['beta1'], h['beta2'], h['eps']
        a.weight_decay, a.momentum, a.grad_scale = float(g.get('weight_decay', 0.0)), h['momentum'], grad_scale
        return a

    def sync_lr(self):
        """Upload the optimizer's current learning rate when it changed (one tiny fill, never inside a capture)."""
        lr = float(self.optimizer.param_groups[0]['lr'])
        if lr != self._lr_uploaded:
            if self.device.type == 'cuda' and _torch.cuda.is_current_stream_capturing():
                raise RuntimeError('learning rate changed during CUDA-graph capture')
            self.lr_dev.fill_(lr)
            self._lr_uploaded = lr

    def _launch(self, world, grad_scale, zero_grads=True):
        nat = self._nat
        variant = self._pick_variant(self.numel * 4) if world > 1 else 'one_shot'
        a = self._fused_args(0, self.numel, variant, world, grad_scale, zero_grads, True)
        nat.check(nat.lib().coinn_fused_reduce_opt(_C.byref(a), 0, nat.stream_ptr(self.device)),
                  'coinn_fused_reduce_opt')
        from .. import ops as _ops
        _ops._count_launch()
        return variant

    def reduce_and_step(self, zero_grads=True):
        """Average gradients over all sites and apply the optimizer.  Returns the variant used."""
        self.steps_done += 1
        self.host_step += 1
        if self.backend == 'nvlink':
            if getattr(self, '_overlap', None) and self._overlap['armed']:
                self._finish_overlap()
                return 'bucketed'
            return self._launch(self.world, 1.0 / self.world, zero_grads)
        if self.backend == 'nccl':
            if self.world > 1:
                _dist.all_reduce(self.flat_grad, op=_dist.ReduceOp.SUM, group=self.group)
            self._launch(1, 1.0 / self.world, zero_grads)
            return 'nccl'
        # ---- torch fallback (CPU / exotic optimizers) ------------------------------------------
        if self.world > 1:
            _dist.all_reduce(self.flat_grad, op=_dist.ReduceOp.SUM, group=self.group)
            self.flat_grad.div_(self.world)
        self.optimizer.step()
        if zero_grads:
            self.flat_grad.zero_()
        return 'torch'

    # ---------------------------------------------------------------- backward overlap (bucketed)
    def enable_overlap(self, bucket_bytes=4 << 20):
        """Launch the fused reduce+update per *bucket* as soon as backward has produced the bucket's last gradient, on a
        side stream, so the cross-GPU exchange of the late layers overlaps the backward pass of the early ones
        (SURVEY §5.8).  Buckets are contiguous arena ranges covering whole parameters, formed in *backward* order (from
        the last parameter towards the first) so the first bucket to complete is a full-sized one.  "Gradient is final"
        arrives either from autograd (``register_post_accumulate_grad_hook``) or, for kernels that accumulate straight
        into ``.grad`` and hand autograd ``None``, from ``ops.linear.notify_grad_written``.  ``reduce_and_step()`` joins
        the side stream and launches the remaining bucket(s) - the last one bumps the step counter.  Works inside a
        CUDA-graph capture (the side stream becomes a parallel branch of the graph).  ``nvlink`` backend only."""
        if self.backend != 'nvlink' or getattr(self, '_overlap', None):
            return self
        from ..ops import linear as _lin
        cap = max(int(bucket_bytes) // 4, 4 * max(self.world, 1))
        ends = [off + _round_up(p.numel(), _ALIGN) for p, off in zip(self.params, self.offsets)]
        buckets, stop, members = [], self.numel, []
        for i in range(len(self.params) - 1, -1, -1):
            members.append(i)
            start = self.offsets[i]
            if stop - start >= cap or i == 0:
                start = 0 if i == 0 else start
                buckets.append({'offset': start, 'numel': stop - start, 'params': members, 'pending': len(members)})
                stop, members = start, []
        assert sum(b['numel'] for b in buckets) == self.numel and ends
        self._overlap = {'buckets': buckets, 'armed': False, 'launched': set(), 'seen': set(),
                         'stream': _torch.cuda.Stream(self.device), 'owner': {}}
        for b_ix, b in enumerate(buckets):
            for i in b['params']:
                self._overlap['owner'][i] = b_ix
        for i, p in enumerate(self.params):
            p.register_post_accumulate_grad_hook(lambda _p, ix=i: self._grad_ready(ix))
            _lin.register_grad_listener(p, lambda ix=i: self._grad_ready(ix))
        return self

    def arm_overlap(self):
        """Call before the LAST micro-batch's backward of a step (earlier micro-batches only accumulate)."""
        ov = getattr(self, '_overlap', None)
        if ov:
            ov['armed'] = True
            ov['launched'] = set()
            ov['seen'] = set()
            for b in ov['buckets']:
                b['pending'] = len(b['params'])

    def _launch_bucket(self, b_ix, last, side=True):
        ov, nat = self._overlap, self._nat
        b = ov['buckets'][b_ix]
        variant = self._pick_variant(b['numel'] * 4) if self.world > 1 else 'one_shot'
        a = self._fused_args(b['offset'], b['numel'], variant, self.world, 1.0 / self.world, True, bool(last))
        cur = _torch.cuda.current_stream(self.device)
        if side:
            ov['stream'].wait_stream(cur)               # the bucket's gradients were produced on the compute stream
            with _torch.cuda.stream(ov['stream']):
                nat.check(nat.lib().coinn_fused_reduce_opt(_C.byref(a), 0, nat.stream_ptr(self.device)),
                          'coinn_fused_reduce_opt[bucket]')
        else:
            nat.check(nat.lib().coinn_fused_reduce_opt(_C.byref(a), 0, nat.stream_ptr(self.device)),
                      'coinn_fused_reduce_opt[bucket]')
        from .. import ops as _ops
        _ops._count_launch()
        ov['launched'].add(b_ix)

    def _grad_ready(self, param_ix):
        ov = self._overlap
        if not ov['armed']:
            return
        if param_ix in ov['seen']:
            return      # one count per parameter and step: autograd runs AccumulateGrad (and its hooks) even for the undefined
        ov['seen'].add(param_ix)   # gradients of direct-mode kernels, which have already reported through notify_grad_written
        b_ix = ov['owner'][param_ix]
        b = ov['buckets'][b_ix]
        b['pending'] -= 1
        if b['pending'] == 0 and len(ov['launched']) < len(ov['buckets']) - 1:
            self._launch_bucket(b_ix, last=False)       # the final bucket is launched by reduce_and_step (bumps the step)

    def _finish_overlap(self):
        """Join the side stream FIRST, then run what is left on the compute stream: launches of one arena never run
        concurrently (they share the flag pad / per-CTA sequence numbers and the device step counter)."""
        ov = self._overlap
        if ov['launched']:      # (never wait on an idle side stream: inside a graph capture that is an isolation error)
            _torch.cuda.current_stream(self.device).wait_stream(ov['stream'])
        rest = [i for i in range(len(ov['buckets'])) if i not in ov['launched']]
        for k, b_ix in enumerate(rest):
            self._launch_bucket(b_ix, last=(k == len(rest) - 1), side=False)
        ov['armed'] = False

    def local_step(self, zero_grads=True):
        """Optimizer step on the local gradients only (pre-training / single site)."""
        self.steps_done += 1
        self.host_step += 1
        if self.backend in ('nvlink', 'nccl'):
            self._launch(1, 1.0, zero_grads)
        else:
            self.optimizer.step()
            if zero_grads:
                self.flat_grad.zero_()


class _Local:
    """Same surface as SymmetricBuffer for buffers that never leave the device."""

    def __init__(self, numel, dtype, device):
        self.local = _torch.zeros(int(numel), dtype=dtype, device=device)
        self.peer_ptrs = [self.local.data_ptr()]
        self.multicast_ptr = 0
        self.handle = None

    def barrier(self):
        pass


class SymmAllReduce:
    """In-kernel all-reduce(mean) of small tensor lists over NVLink (no NCCL): the PowerSGD P / Q factor exchange
    and the dense part of rankDAD.  Same kernel as the dSGD step with ``opt_kind = NONE``:
    every rank's values are packed into a symmetric input buffer, ``fused_reduce_opt_kernel`` writes the mean into
    the symmetric output buffer of every rank (one-shot: redundantly; two-shot / NVLS: shard + broadcast) and
    re-zeroes the input.  Falls back to ``torch.distributed.all_reduce`` off-GPU."""

    def __init__(self, capacity, device, group=None, variant='auto'):
        self.device, self.group = _torch.device(device), group
        self.world, self.rank = _world(group), _rank(group)
        self.capacity = _round_up(max(int(capacity), 4), 4 * max(self.world, 1))
        self.variant = variant
        self.native = self.device.type == 'cuda'
        if self.native:
            from ..ops import native as _nat
            self._nat = _nat
            mk = (lambda n, dt: SymmetricBuffer(n, dt, self.device, group)) if self.world > 1 else \
                (lambda n, dt: _Local(n, dt, self.device))
            self.inp, self.out = mk(self.capacity, _torch.float32), mk(self.capacity, _torch.float32)
            self.flags = mk(_nat.lib().coinn_fused_flag_slots(), _torch.int32)
            self.epoch = _torch.zeros(_nat.lib().coinn_fused_max_blocks(), dtype=_torch.int32, device=self.device)
            self.ticket = _torch.zeros(1, dtype=_torch.int32, device=self.device)
            self.step = _torch.zeros(1, dtype=_torch.int32, device=self.device)
            self.error = _torch.zeros(1, dtype=_torch.int32, device=self.device)
            self.dummy = _torch.zeros(4, dtype=_torch.float32, device=self.device)

    def mean_inplace(self, numel):
        """Zero-copy form: the caller has written ``inp.local[:numel]`` (e.g. a kernel produced the PowerSGD factors
        right there); afterwards ``out.local[:numel]`` holds the mean over all ranks and ``inp.local`` is zero again.
        Returns ``out.local``.  Native only (CUDA)."""
        assert self.native and numel <= self.capacity
        if self.world == 1:
            self.out.local[:numel].copy_(self.inp.local[:numel])
            self.inp.local[:numel].zero_()
            return self.out.local
        self._launch_mean(_round_up(int(numel), 4 * self.world))
        return self.out.local

    def _launch_mean(self, n4):
        nat = self._nat
        v = self.variant
        if v == 'auto':
            v = 'one_shot' if n4 * 4 <= _conf.ONE_SHOT_MAX_BYTES else ('nvls' if self.inp.multicast_ptr and self.out.multicast_ptr else 'two_shot')
        a = nat.FusedArgs()
        for r in range(self.world):
            a.grad_ptrs[r], a.param_ptrs[r], a.flag_ptrs[r] = self.inp.peer_ptrs[r], self.out.peer_ptrs[r], self.flags.peer_ptrs[r]
        a.grad_mc = self.inp.multicast_ptr or None
        a.param_mc = self.out.multicast_ptr or None
        a.m, a.v = self.dummy.data_ptr(), self.dummy.data_ptr()
        a.epoch, a.step, a.ticket = self.epoch.data_ptr(), self.step.data_ptr(), self.ticket.data_ptr()
        a.offset, a.numel = 0, n4
        a.rank, a.world, a.variant, a.opt_kind, a.grad_dtype = self.rank, self.world, _VARIANTS[v], _OPT_KINDS['none'], 0
        a.zero_grads, a.bump_step, a.grad_scale = 1, 0, 1.0 / self.world
        a.error, a.timeout_ms = self.error.data_ptr(), _DEFAULT_TIMEOUT_MS
        nat.check(nat.lib().coinn_fused_reduce_opt(_C.byref(a), 0, nat.stream_ptr(self.device)), 'fused all-reduce')
        from .. import ops as _ops
        _ops._count_launch()

    def mean_(self, tensors):
        """Replace every tensor in ``tensors`` by its mean over all ranks (in place)."""
        tensors = [t for t in tensors if t.numel()]
        if not tensors or self.world == 1:
            return tensors
        total = sum(t.numel() for t in tensors)
        if not self.native or total > self.capacity:
            flat = _torch.cat([t.reshape(-1).float() for t in tensors])
            _dist.all_reduce(flat, group=self.group)
            flat /= self.world
        else:
            off = 0
            for t in tensors:
                self.inp.local[off:off + t.numel()].copy_(t.reshape(-1))
                off += t.numel()
            self._launch_mean(_round_up(total, 4 * self.world))
            flat = self.out.local
        off = 0
        for t in tensors:
            t.copy_(flat[off:off + t.numel()].view_as(t))
            off += t.numel()
        return tensors


class SymmAllGather:
    """All-gather of one flat fp32 buffer per site over symmetric memory (``lowrank.cu::allgather_kernel``): every site
    writes its payload into its own peer-mapped send buffer, one launch copies all sites' payloads into a local
    ``[world, numel]`` buffer (16-byte peer loads over NVLink between two flag barriers).  The rankDAD factor exchange
    (C4 / K12); no NCCL.  ``torch.distributed.all_gather`` off-GPU."""

    def __init__(self, capacity, device, group=None):
        self.device, self.group = _torch.device(device), group
        self.world, self.rank = _world(group), _rank(group)
        self.capacity = _round_up(max(int(capacity), 4), 4)
        self.native = self.device.type == 'cuda'
        if self.native and self.world > 1:
            from ..ops import native as _nat
            self._nat = _nat
            self.send = SymmetricBuffer(self.capacity, _torch.float32, self.device, group)
            self.flags = SymmetricBuffer(_nat.lib().coinn_allgather_flag_slots(), _torch.int32, self.device, group)
            self.epoch = _torch.zeros(_nat.lib().coinn_allgather_max_blocks(), dtype=_torch.int32, device=self.device)
            self.error = _torch.zeros(1, dtype=_torch.int32, device=self.device)
        else:
            self.send = _Local(self.capacity, _torch.float32, self.device)
        self.recv = _torch.zeros(max(self.world, 1) * self.capacity, dtype=_torch.float32, device=self.device)

    def gather(self, numel):
        """``send.local[:numel]`` of every site -> returns ``recv`` viewed as ``[world, stride]`` (row s = site s's
        payload in ``[:numel]``) and the row stride in elements."""
        n4 = _round_up(int(numel), 4)
        assert n4 <= self.capacity
        if self.world == 1:
            self.recv[:n4].copy_(self.send.local[:n4])
            return self.recv[:n4].view(1, n4), n4
        if not self.native:
            parts = [_torch.empty(n4, dtype=_torch.float32, device=self.device) for _ in range(self.world)]
            _dist.all_gather(parts, self.send.local[:n4].contiguous(), group=self.group)
            self.recv[:self.world * n4].copy_(_torch.cat(parts))
            return self.recv[:self.world * n4].view(self.world, n4), n4
        nat = self._nat
        a = nat.AllGatherArgs()
        for r in range(self.world):
            a.src_ptrs[r], a.flag_ptrs[r] = self.send.peer_ptrs[r], self.flags.peer_ptrs[r]
        a.dst, a.epoch, a.error = self.recv.data_ptr(), self.epoch.data_ptr(), self.error.data_ptr()
        a.numel, a.rank, a.world, a.timeout_ms = n4, self.rank, self.world, _DEFAULT_TIMEOUT_MS
        nat.check(nat.lib().coinn_allgather(_C.byref(a), nat.stream_ptr(self.device)), 'coinn_allgather')
        from .. import ops as _ops
        _ops._count_launch()
        return self.recv[:self.world * n4].view(self.world, n4), n4
