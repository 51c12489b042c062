"""Learners whose per-step gradient exchange happens inside the fused sm_100a kernel.

``NvlinkLearner`` keeps the COINNLearner surface (``step`` / ``backward`` / ``to_reduce``) but one
``to_reduce()`` call runs *a whole local epoch*: every step is ``backward`` (``local_iterations``
micro-batches, gradients accumulate straight into the symmetric arena) followed by
``arena.reduce_and_step()`` - gradient mean over all sites + optimizer update in one kernel, no
NCCL, no host copy, no JSON.  The JSON control plane is only touched when the epoch is over
(``mode = validation_waiting``), i.e. once per epoch instead of once per step.

Epoch length: all sites must launch the same number of fused steps, so the number of steps per
round is ``max`` over sites of their local step count (one tiny all-reduce per epoch).  Sites with
less data wrap around and keep contributing - the reference's "lagging sites re-shuffle and
continue until everyone is waiting" rule (local.py:232-238), decided up-front instead of round
by round.

The same code runs on CPU with the gloo backend (``DistArena`` falls back to all-reduce +
``optimizer.step``), which is how the host-side logic is tested without a GPU.
"""
import time as _time

import torch as _torch
import torch.distributed as _dist

from ..config.keys import Mode, Transport
from ..distrib.learner import COINNLearner
from .arena import DistArena, wire_dtype_for


def _dist_on():
    return _dist.is_available() and _dist.is_initialized()


class _LaggedReadback:
    """Per-step device->host read of the loss without stalling the launch pipeline: step t's loss is copied
    asynchronously into a slot of a pinned ring and *consumed* (event-synchronised) ``lag`` steps later, so the host
    runs up to ``lag`` steps ahead of the device and a scheduling hiccup of a few milliseconds never drains the GPU
    queue.  The ring is allocated once per site (``cache['_readback']``) - pinning host memory inside the round costs
    a driver call that serialises against every other rank of the box."""

    def __init__(self, device, slots=8, lag=4):
        self.cuda = device.type == 'cuda'
        self.buf = _torch.zeros(slots, dtype=_torch.float32)
        if self.cuda:
            self.buf = self.buf.pin_memory()
        self.events = [_torch.cuda.Event() for _ in range(slots)] if self.cuda else [None] * slots
        self.slots, self.lag = slots, min(lag, slots - 1)
        self.reset()

    def reset(self):
        self.pending = [False] * self.slots
        self.count, self.head, self.last = 0, 0, float('nan')
        return self

    @classmethod
    def for_site(cls, cache, device):
        rb = cache.get('_readback')
        if rb is None or rb.cuda != (device.type == 'cuda'):
            rb = cache['_readback'] = cls(device, slots=int(cache.get('readback_slots', 8)),
                                          lag=int(cache.get('readback_lag', 4)))
        return rb.reset()

    def _consume(self, i):
        if self.pending[i]:
            if self.cuda:
                self.events[i].synchronize()
            self.last = float(self.buf[i])
            self.pending[i] = False
            self.count += 1

    def push(self, loss):
        i = self.head % self.slots
        self._consume(i)                                   # slot reuse bounds the lag even if `lag` >= slots
        self.buf[i:i + 1].copy_(loss.detach().reshape(1), non_blocking=True)
        if self.cuda:
            self.events[i].record()
        self.pending[i] = True
        if self.head >= self.lag:
            self._consume((self.head - self.lag) % self.slots)   # read step t-lag now that step t is queued
        self.head += 1

    def finish(self):
        for k in range(self.slots):
            self._consume((self.head + k) % self.slots)
        return self.last


class _RunningScores:
    """Folds every step's ``averages`` / ``metrics`` into two accumulators so an epoch of fused steps keeps O(1)
    iteration dicts alive instead of O(steps) (loss / logits / prediction tensors of every step)."""

    def __init__(self, trainer):
        self.trainer = trainer
        self.avg, self.met, self.last = trainer.new_averages(), trainer.new_metrics(), None

    def add(self, its):
        for it in its:
            if it.get('averages') is not None:
                self.avg.accumulate(it['averages'])
            if it.get('metrics') is not None:
                self.met.accumulate(it['metrics'])
        if its:
            self.last = its[-1]

    def result(self):
        out = dict(self.trainer.reduce_iteration([self.last])) if self.last is not None else {}
        out['averages'], out['metrics'] = self.avg, self.met
        return out


class NvlinkLearner(COINNLearner):
    _overlap_ok = True      # bucket kernels from grad hooks only make sense when reduce+update is the fused kernel
    @property
    def arena(self):
        """The flat arenas + fused optimizer of this site; built on first use (the learner object
        is re-created every round, also in phases where no model exists yet)."""
        arena = self.cache.get('_arena')
        if arena is None or arena.model is not self.model or arena.optimizer is not self.optim:
            transport = self.cache.get('transport', Transport.NVLINK)
            backend = {'nvlink': 'nvlink', 'nccl': 'nccl'}.get(str(transport), 'auto')
            if self.device.type != 'cuda':
                backend = 'torch'
            arena = DistArena(self.model, self.optim, device=self.device, backend=backend,
                              variant=self.cache.get('reduce_variant', 'auto'),
                              shadow_bf16=bool(self.cache.get('shadow_bf16', False)),
                              grad_dtype=wire_dtype_for(self.cache),
                              timeout_ms=(float(self.cache['barrier_timeout_s']) * 1e3
                                          if self.cache.get('barrier_timeout_s') else None))
            if self.cache.get('overlap_backward') and self._overlap_ok:
                arena.enable_overlap(int(self.cache.get('bucket_bytes', 4 << 20)))
            self.cache['_arena'] = arena
        return arena

    # the update is fused into to_reduce(); nothing left to apply when the remote says `update`
    def step(self) -> dict:
        return {}

    def _steps_this_round(self):
        """max over sites of local steps per epoch (so every site launches the same kernels)."""
        fixed = self.cache.get('steps_per_round')
        if fixed:
            return int(fixed)
        ds = self.trainer.data_handle.dataset.get('train')
        bs = int(self.cache['batch_size']) * int(self.cache.get('local_iterations', 1))
        mine = -(-len(ds) // bs) if ds is not None and len(ds) else 1
        if _dist_on() and _dist.get_world_size() > 1:
            t = _torch.tensor([mine], dtype=_torch.int64, device=self.device if self.device.type == 'cuda' else 'cpu')
            _dist.all_reduce(t, op=_dist.ReduceOp.MAX)
            mine = int(t.item())
        return max(mine, 1)

    def backward(self):
        """Micro-batches accumulate into the arena (it is zeroed by the fused kernel's tail)."""
        out, its = {}, []
        self.model.train()
        self.arena.rebind_grads()
        k = self.cache.get('local_iterations', 1)
        for i in range(k):
            batch, flags = self.trainer.data_handle.next_iter()
            it = self.trainer.iteration(batch)
            if i == k - 1 and self._overlap_ok:
                self.arena.arm_overlap()               # bucket kernels may start as soon as their gradients are final
            it['loss'].backward()
            its.append(it)
            out.update(**flags)
        return its, out

    def _end_of_round_cursor(self):
        """The round IS the epoch: report it finished regardless of where the local cursor is and start the next round
        with a fresh (re-shuffled) loader.  With an explicit ``steps_per_round`` the site streams instead: the loader
        iterator - and the batches its prefetcher has already staged on the device - carry over into the next round."""
        if not self.cache.get('steps_per_round'):
            self.cache['cursor'] = 0

    def _graphed_round(self, steps):
        """``cache['cuda_graph']``: capture the whole step once, then replay it ``steps`` times."""
        from .graph_step import GraphedStep
        gs = self.cache.get('_graph_step')
        done = 0
        if gs is None or gs.arena is not self.arena:
            batch, _ = self.trainer.data_handle.next_iter()
            gs = self.cache['_graph_step'] = GraphedStep(self).capture(batch)   # state is rolled back after warm-up
            gs.step(batch)                                                       # ... so this is training step 1
            done = 1
        rb = _LaggedReadback.for_site(self.cache, self.device) if self.cache.get('readback_per_step') else None
        if rb is not None and done:
            rb.push(gs.it['loss'])
        stamps = self.cache.get('_host_stamps')             # bench/diagnostics: host time at which each step was queued
        for _ in range(steps - done):
            batch, _ = self.trainer.data_handle.next_iter()
            gs.step(batch)
            if rb is not None:
                rb.push(gs.it['loss'])
            if stamps is not None:
                stamps.append(_time.perf_counter())
        if rb is not None:
            self.cache['last_loss'] = rb.finish()
            self.cache['losses_read'] = rb.count
        avg, met = gs.drain()
        return {'averages': avg, 'metrics': met}

    def to_reduce(self):
        its, out = [], {}
        graphable = (self.cache.get('cuda_graph') and self.device.type == 'cuda'
                     and self.cache.get('local_iterations', 1) == 1 and self.arena.backend == 'nvlink')
        if graphable:
            it = self._graphed_round(self._steps_this_round())
            self.arena.check_health()                     # barrier watchdog: a site that never arrived raises here
            self._end_of_round_cursor()
            out['mode'] = Mode.VALIDATION_WAITING
            out['fused_steps'] = self.arena.steps_done
            return it, out
        rb = _LaggedReadback.for_site(self.cache, self.device) if self.cache.get('readback_per_step') else None
        timer = None
        if self.cache.get('profile'):                     # compspec-style "profile": true -> per-phase device timers
            from ..utils.profiling import DeviceTimer
            timer = self.cache.setdefault('_profile_timer', DeviceTimer(self.device if self.device.type == 'cuda' else None))
        scores = _RunningScores(self.trainer)
        for _ in range(self._steps_this_round()):
            if timer is not None:
                with timer('forward_backward'):
                    step_its, flags = self.backward()
                with timer('reduce_update'):
                    self.arena.reduce_and_step()
            else:
                step_its, flags = self.backward()
                self.arena.reduce_and_step()
            if rb is not None:                            # end-to-end mode: every step's loss goes to the host
                rb.push(step_its[-1]['loss'])
            scores.add(step_its)
        if rb is not None:
            self.cache['last_loss'] = rb.finish()
            self.cache['losses_read'] = rb.count
        if timer is not None:
            from ..utils.profiling import site_profile
            out['profile'] = site_profile(self.cache, timer, key='profile_log')
        self.arena.check_health()
        self._end_of_round_cursor()
        out['mode'] = Mode.VALIDATION_WAITING
        out['fused_steps'] = self.arena.steps_done
        return scores.result(), out


class NvlinkPowerSGDLearner(NvlinkLearner):
    """PowerSGD over the device collectives (P/Q factors all-reduced instead of shipped as files).
    Math as in ``distrib.powersgd`` (rank-r, error feedback, warm start).  On GPUs a compressed step is
    ``_compressed_step_device`` (batched kernels + in-kernel NVLink all-reduce of the factor buffers); the per-matrix
    PyTorch form below it is the CPU / gloo path and the oracle of the device path."""
    _overlap_ok = False

    def __init__(self, **kw):
        super().__init__(**kw)
        from ..distrib.powersgd import PowerSGDState
        c = self.cache
        self.rank_r = c.setdefault('matrix_approximation_rank', 1)
        self.start_iter = c.setdefault('start_powerSGD_iter', 10)
        self.error_feedback = c.setdefault('use_error_feedback', True)
        self.warm_start = c.setdefault('warm_start', True)
        self.seed = c.get('seed') or 0
        self.st = c.setdefault('powerSGD_state', PowerSGDState())

    def _allreduce_mean(self, tensors):
        """P / Q factors and rank-1 gradients: in-kernel NVLink all-reduce on GPUs (SymmAllReduce), gloo/NCCL
        all-reduce elsewhere."""
        if not tensors or not (_dist_on() and _dist.get_world_size() > 1):
            return
        red = self.cache.get('_symm_allreduce')
        need = sum(t.numel() for t in tensors)
        if red is None or red.capacity < need:
            from .arena import SymmAllReduce
            red = self.cache['_symm_allreduce'] = SymmAllReduce(max(need, 1 << 16), self.device)
        red.mean_(tensors)

    def _device_path(self):
        return self.device.type == 'cuda' and self.arena.backend in ('nvlink', 'nccl') \
            and int(self.rank_r) <= 8 and self.cache.get('lowrank_kernels', True)

    def _compressed_step_device(self):
        """One compressed step entirely on the hand-written kernels (``ops/lowrank.py`` + the fused all-reduce): no
        ``torch.distributed`` collective, no ``torch.matmul``.  Ten launches for the whole model:

            orthogonalize(Q) | M = G + E, P = M Q | all-reduce(P) | orthogonalize(P) | Q = M^T P | gather rank-1 grads
            | all-reduce(Q + rank-1) | scatter rank-1 | G = P Q^T, E = M - G | fused local optimizer step

        The factor buffers ARE the symmetric exchange buffers (kernels write P / Q straight into ``SymmAllReduce.inp`` and
        read the averaged factors from ``.out``), the warm-start Q is the averaged Q of the previous step, and the error
        memory is one flat buffer laid out like the gradient arena (ref powersgd/__init__.py:61-181)."""
        from ..ops.lowrank import PowerSGDPlan
        from .arena import SymmAllReduce
        arena, st, c = self.arena, self.st, self.cache
        plan = c.get('_psgd_plan')
        if plan is None or plan.rank != int(self.rank_r) or c.get('_psgd_arena') is not arena:
            plan = c['_psgd_plan'] = PowerSGDPlan(arena.params, arena.offsets, int(self.rank_r), self.device)
            c['_psgd_arena'] = arena
            c['_psgd_error'] = _torch.zeros(arena.numel, dtype=_torch.float32, device=self.device)
            c['_psgd_xp'] = SymmAllReduce(max(plan.p_numel, 4), self.device)
            c['_psgd_xq'] = SymmAllReduce(max(plan.q_numel + plan.low_numel, 4), self.device)
            c['_psgd_q_ready'] = False
        E, xp, xq = c['_psgd_error'], c['_psgd_xp'], c['_psgd_xq']
        Qavg = xq.out.local                                       # averaged Q of the previous step (warm start)
        if not (self.warm_start and c['_psgd_q_ready']):
            gen = _torch.Generator(device='cpu').manual_seed(int(self.seed) + int(st.iter))   # same Q on every site
            Qavg[:plan.q_numel].copy_(_torch.randn(plan.q_numel, generator=gen).to(self.device))
            c['_psgd_q_ready'] = True
        G = arena.flat_grad
        plan.orthogonalize(Qavg, which=1)
        plan.mq(G, E, Qavg, xp.inp.local, self.error_feedback)    # E := M = G + E ; P -> exchange buffer
        Pavg = xp.mean_inplace(plan.p_numel)
        plan.orthogonalize(Pavg, which=0)
        plan.mtp(E, Pavg, xq.inp.local)                           # Q = M^T P -> exchange buffer
        plan.gather_low(G, xq.inp.local)                          # biases / norms ride along uncompressed
        Qavg = xq.mean_inplace(plan.q_numel + plan.low_numel)
        plan.scatter_low(Qavg, G)
        plan.reconstruct(G, E, Pavg, Qavg, self.error_feedback)   # G = P Q^T ; E = M - G
        arena.local_step()

    def _compressed_step(self):
        if self._device_path():
            return self._compressed_step_device()
        from ..distrib.powersgd import _native_orthogonalize, _as_matrix
        st = self.st
        mats, low = [], []
        for key, p in self.model.named_parameters():
            if p.grad is None:
                continue
            (low if p.dim() <= 1 else mats).append((key, p))
        Ps = []
        for key, p in mats:
            M = _as_matrix(p.grad.detach().float().clone())
            if self.error_feedback and key in st.error_dict:
                M += st.error_dict[key]
            if not self.warm_start or key not in st.q_memory_dict:
                gen = _torch.Generator(device='cpu').manual_seed(int(self.seed) + int(st.iter))
                st.q_memory_dict[key] = _torch.randn(M.shape[1], self.rank_r, generator=gen).to(M.device)
            _native_orthogonalize(st.q_memory_dict[key])
            st.high_rank_tensors[key] = M
            st.p_memory_dict[key] = M @ st.q_memory_dict[key]
            Ps.append(st.p_memory_dict[key])
        self._allreduce_mean(Ps)
        Qs = []
        for key, p in mats:
            _native_orthogonalize(st.p_memory_dict[key])
            st.q_memory_dict[key] = st.high_rank_tensors[key].t() @ st.p_memory_dict[key]
            Qs.append(st.q_memory_dict[key])
        self._allreduce_mean(Qs + [p.grad for _, p in low])
        for key, p in mats:
            approx = st.p_memory_dict[key] @ st.q_memory_dict[key].t()
            if self.error_feedback:
                st.error_dict[key] = st.high_rank_tensors[key] - approx
            p.grad.copy_(approx.view_as(p.grad))
        st.high_rank_tensors.clear()
        self.arena.local_step()

    def to_reduce(self):
        out, scores = {}, _RunningScores(self.trainer)
        for _ in range(self._steps_this_round()):
            step_its, _ = self.backward()
            if self.st.iter < self.start_iter:
                self.arena.reduce_and_step()
            else:
                if self.st.iter == self.start_iter:
                    # warm-up ran sharded updates (two-shot / NVLS: rank r holds the Adam moments of shard r only);
                    # the compressed phase updates the full range locally, so every site needs the complete moments
                    self.arena.gather_state()
                self._compressed_step()
            self.st.iter += 1
            scores.add(step_its)
        self.arena.check_health()
        self.cache['cursor'] = 0
        out['mode'] = Mode.VALIDATION_WAITING
        return scores.result(), out


class NvlinkDADLearner(NvlinkLearner):
    """rankDAD over device collectives: per-layer (delta, activation) factors are all-gathered,
    concatenated along the rank axis and re-compressed on every site (deterministically, so the
    replicas agree), then turned back into dense gradients for the fused local step."""
    _overlap_ok = False

    def __init__(self, **kw):
        from ..distrib.rankdad.spi import DADParallel
        COINNLearner.__init__(self, **kw)
        for key in list(self.trainer.nn):
            if not isinstance(self.trainer.nn[key], DADParallel):
                self.trainer.nn[key] = DADParallel(self.trainer.nn[key], cache=self.cache, input=self.input,
                                                   state=self.state, device=self.trainer.device['gpu'],
                                                   dtype=self.dtype)
        super().__init__(**kw)

    def _dad_step_device(self):
        """rankDAD on the device data plane (C4 / K11 / K12): per layer the local (delta, activation) factors come from
        the coefficient-space power iteration (``ops.lowrank.lowrank_factor``: two Gram launches, one single-CTA
        eigen-solver, two skinny GEMMs - no cuSOLVER, no host sync) and are written straight into the symmetric send
        buffer; ONE all-gather launch (peer loads over NVLink) collects every site's factors of every layer; each layer is
        re-compressed from the gathered column blocks in place and ``dad_reconstruct`` writes ``delta act^T`` (+ the bias
        column) directly into the gradient arena.  Dense (non-DAD) parameters go through the fused all-reduce.  Zero
        ``torch.distributed`` collectives, zero ``torch.matmul`` (ref rankdad/__init__.py:63-98, spi.py:190-250)."""
        from ..distrib.rankdad.spi import _mm_flatten
        from ..ops.lowrank import dad_reconstruct, lowrank_factor
        from .arena import SymmAllGather, SymmAllReduce
        wrapper, c = self.model, self.cache
        world = _dist.get_world_size() if _dist_on() else 1
        rank_r, iters, tol = int(wrapper.rank), int(wrapper.num_pow_iters), float(wrapper.dad_tol)
        layers = wrapper.dad_layers(reverse=True)
        # ---- layout of the exchange buffer: per layer [left (out x k) | right (in(+1) x k)], 4-element aligned
        shapes, total = [], 0
        for name, m in layers:
            delta, act = _mm_flatten(wrapper._local_grads[name].float(), wrapper._activations[name].float())
            rows_c = act.shape[1] + (1 if (wrapper.bias_augment and getattr(m, 'bias', None) is not None) else 0)
            k = max(1, min(rank_r, delta.shape[1], rows_c, delta.shape[0]))
            lo, ro = total, total + -(-delta.shape[1] * k // 4) * 4
            total = ro + -(-rows_c * k // 4) * 4
            shapes.append((name, m, delta, act, rows_c, k, lo, ro))
        ag = c.get('_dad_gather')
        if ag is None or ag.capacity < total:
            ag = c['_dad_gather'] = SymmAllGather(max(total, 4), self.device)
        send = ag.send.local
        for name, m, delta, act, rows_c, k, lo, ro in shapes:
            if rows_c == act.shape[1] + 1:
                act = _torch.cat([act, act.new_ones(act.shape[0], 1)], dim=1)
            lowrank_factor(delta.t().contiguous(), act.t().contiguous(), rank_r, iters, tol,
                           out_left=send[lo:lo + delta.shape[1] * k].view(delta.shape[1], k),
                           out_right=send[ro:ro + rows_c * k].view(rows_c, k))
        recv, stride = ag.gather(total)                            # [world, stride]
        scale = (1.0 / world) if c.get('dad_mean') else 1.0
        for name, m, delta, act, rows_c, k, lo, ro in shapes:
            out_f = delta.shape[1]
            if world > 1 and world * k > rank_r and c.get('dad_recompress', True) and world * k <= 96:
                left, right = lowrank_factor(None, None, rank_r, iters, tol,
                                             b_seg=(recv[0, lo:], out_f, world * k, k, stride),
                                             c_seg=(recv[0, ro:], rows_c, world * k, k, stride))
            elif world > 1:                                        # no re-compression: concatenate the column blocks
                left = _torch.cat([recv[s, lo:lo + out_f * k].view(out_f, k) for s in range(world)], 1)
                right = _torch.cat([recv[s, ro:ro + rows_c * k].view(rows_c, k) for s in range(world)], 1)
            else:
                left, right = recv[0, lo:lo + out_f * k].view(out_f, k), recv[0, ro:ro + rows_c * k].view(rows_c, k)
            has_bias = getattr(m, 'bias', None) is not None
            if left.shape[1] <= 16:
                dad_reconstruct(left, right, m.weight.grad, m.bias.grad if has_bias else None, scale=scale)
                if has_bias and rows_c == m.weight.shape[1]:       # no bias column carried: the reference's approximation
                    m.bias.grad.copy_(left.sum(1) * scale)
            else:                                                  # wide un-recompressed factors: plain product
                full = (left * scale) @ right.t()
                m.weight.grad.copy_(full[:, :m.weight.shape[1]])
                if has_bias:
                    m.bias.grad.copy_(full[:, -1] if rows_c == m.weight.shape[1] + 1 else left.sum(1) * scale)
        plain = [p.grad for p in wrapper.plain_parameters() if p.grad is not None]
        if plain and world > 1:
            red = c.get('_symm_allreduce')
            need = sum(t.numel() for t in plain)
            if red is None or red.capacity < need:
                red = c['_symm_allreduce'] = SymmAllReduce(max(need, 1 << 16), self.device)
            red.mean_(plain)
            if not c.get('dad_mean'):
                _torch._foreach_mul_(plain, float(world))          # rankDAD exchanges sums (quirk 8.5-9)
        self.arena.local_step()

    def _dad_step(self):
        if self.device.type == 'cuda' and self.arena.backend in ('nvlink', 'nccl') and self.cache.get('lowrank_kernels', True) \
                and int(self.model.rank) <= 16:
            return self._dad_step_device()
        from ..distrib.rankdad.spi import power_iteration_BC
        wrapper = self.model
        world = _dist.get_world_size() if _dist_on() else 1
        rank_r, iters, tol = wrapper.rank, wrapper.num_pow_iters, wrapper.dad_tol
        for name, m in wrapper.dad_layers(reverse=True):
            delta, act = wrapper._factors(name, m)                # [out,k], [in(+1),k]
            if world > 1:
                ds = [_torch.empty_like(delta) for _ in range(world)]
                as_ = [_torch.empty_like(act) for _ in range(world)]
                _dist.all_gather(ds, delta.contiguous())
                _dist.all_gather(as_, act.contiguous())
                delta, act = _torch.cat(ds, 1), _torch.cat(as_, 1)
                if delta.shape[1] > rank_r and self.cache.get('dad_recompress', True):
                    delta, act = power_iteration_BC(delta, act, rank_r, iters, tol)
            if self.cache.get('dad_mean'):
                delta = delta / world
            full = delta @ act.t()
            if getattr(m, 'bias', None) is not None and act.shape[0] == m.weight.shape[1] + 1:
                m.weight.grad.copy_(full[:, :-1])
                m.bias.grad.copy_(full[:, -1])
            else:
                m.weight.grad.copy_(full)
        plain = [p.grad for p in wrapper.plain_parameters() if p.grad is not None]
        if plain and world > 1:
            flat = _torch.cat([g.reshape(-1) for g in plain])
            _dist.all_reduce(flat)
            if self.cache.get('dad_mean'):
                flat /= world
            off = 0
            for g in plain:
                g.copy_(flat[off:off + g.numel()].view_as(g))
                off += g.numel()
        self.arena.local_step()

    def to_reduce(self):
        out, scores = {}, _RunningScores(self.trainer)
        saved = self.cache.get('local_iterations', 1)
        self.cache['local_iterations'] = 1          # rankDAD cannot accumulate gradients
        try:
            for _ in range(self._steps_this_round()):
                step_its, _ = self.backward()
                self._dad_step()
                scores.add(step_its)
        finally:
            self.cache['local_iterations'] = saved
        self.arena.check_health()
        self.cache['cursor'] = 0
        out['mode'] = Mode.VALIDATION_WAITING
        return scores.result(), out
