"""Whole-step CUDA graph: forward + loss + metrics + backward + fused reduce/optimizer in ONE replay.

The reference pays Python + launch latency for every tiny op of a step and several ``.item()`` syncs
(SURVEY §2.5 K5/K8, §7.3-4); for the 61 k-parameter FreeSurfer MLP that *is* the step time.  Here the
user's ``trainer.iteration(batch)`` is captured once - with whatever metric objects it creates - and then
replayed; per-step scores are pushed, inside the graph, into a device ring buffer and read back once per
epoch (``drain``), so a training round performs zero host syncs.

Requirements: static batch shapes (the padded sampler guarantees them), ``local_iterations == 1``, the built-in
metric objects (COINNAverages / Prf1a / ConfusionMatrix keep their per-step counters in device tensors).
"""
import torch as _torch

from ..metrics import COINNAverages, ConfusionMatrix, Prf1a


def _flatten_batch(batch):
    if isinstance(batch, dict):
        return {k: v for k, v in batch.items() if isinstance(v, _torch.Tensor)}
    raise TypeError('GraphedStep needs dict batches of tensors')


def _state_tensors(avg, met):
    """Device tensors carrying one step's scores (all float64 views are made at drain time)."""
    parts = []
    if isinstance(avg, COINNAverages):
        parts += [t.reshape(1).float() for _, t in avg._pending]
    if isinstance(met, (Prf1a, ConfusionMatrix)) and met._dev is not None:
        parts.append(met._dev.reshape(-1).float())
    return parts


class GraphedStep:
    def __init__(self, learner, ring_len=4096):
        self.learner = learner
        self.trainer = learner.trainer
        self.arena = learner.arena
        self.device = learner.device
        self.ring_len = ring_len
        self.graph = None
        self.static = None
        self.it = None
        self.ring = None
        self.cursor = None           # device-side write index
        self.steps_in_ring = 0
        self.kernels_per_replay = 0

    # --------------------------------------------------------------------------------- capture
    def _one_step(self):
        it = self.trainer.iteration(self.static)
        self.arena.arm_overlap()
        it['loss'].backward()
        self.arena.reduce_and_step()
        return it

    def capture(self, batch):
        from .. import ops as _ops
        host = _flatten_batch(batch)
        self.static = {k: _torch.empty_like(v, device=self.device) for k, v in host.items()}
        for k, v in host.items():
            self.static[k].copy_(v, non_blocking=True)
        self.learner.model.train()
        self.arena.rebind_grads()
        # the warm-up passes below are real optimizer steps; snapshot everything they mutate and roll back
        # afterwards so that the first replay is exactly the first training step of the eager schedule
        ar = self.arena
        snap = [t.clone() for t in (ar.flat_param, ar.m, ar.v, ar.step_count)]
        bufs = [(b, b.clone()) for b in self.learner.model.buffers()]
        counters = (ar.steps_done, ar.host_step)
        side = _torch.cuda.Stream(self.device)
        side.wait_stream(_torch.cuda.current_stream(self.device))
        with _torch.cuda.stream(side):                      # warm-up: lazy inits, cudaFuncSetAttribute, autotune
            for _ in range(2):
                warm = self._one_step()
            # the score ring must live OUTSIDE the graph's private pool: a zeros() captured inside the graph
            # would be re-executed (re-zeroing every row) on each replay
            width = sum(t.numel() for t in _state_tensors(warm['averages'], warm['metrics'])) or 1
            if self.ring is None or self.ring.shape[1] != width:
                self.ring = _torch.zeros(self.ring_len, width, dtype=_torch.float32, device=self.device)
        # Drop the warm-up autograd graph BEFORE capturing.  While `warm['loss']` is alive the parameters' AccumulateGrad
        # nodes created on the warm-up stream are reused by the captured backward; a node that receives no gradient
        # (kernels that accumulate straight into .grad return None) then makes the engine wait on that un-captured stream
        # at the end of backward: cudaErrorStreamCaptureIsolation.
        warm = None
        import gc
        gc.collect()
        _torch.cuda.current_stream(self.device).wait_stream(side)
        _torch.cuda.synchronize(self.device)

        self.cursor = _torch.zeros(1, dtype=_torch.int64, device=self.device)
        self.graph = _torch.cuda.CUDAGraph()
        before = _ops.launch_count
        with _torch.cuda.graph(self.graph):
            it = self._one_step()
            parts = _state_tensors(it['averages'], it['metrics'])
            row = _torch.cat(parts) if parts else _torch.zeros(1, device=self.device)
            self.ring.index_copy_(0, self.cursor % self.ring_len, row.unsqueeze(0))
            self.cursor += 1
        self.kernels_per_replay = _ops.launch_count - before
        self.it = it
        self.steps_in_ring = 0
        self.cursor.zero_()
        with _torch.no_grad():
            for dst, src in zip((ar.flat_param, ar.m, ar.v, ar.step_count), snap):
                dst.copy_(src)
            for b, saved in bufs:
                b.copy_(saved)
            ar.flat_grad.zero_()
            if ar.shadow_buf is not None:
                ar.shadow_buf.local.copy_(ar.flat_param)
        ar.steps_done, ar.host_step = counters
        return self

    # ---------------------------------------------------------------------------------- replay
    def step(self, batch):
        from .. import ops as _ops
        for k, v in _flatten_batch(batch).items():
            self.static[k].copy_(v, non_blocking=True)
        self.arena.sync_lr()            # the graph reads the learning rate from a device scalar: schedulers keep working
        self.graph.replay()
        self.arena.steps_done += 1
        self.arena.host_step += 1
        self.steps_in_ring += 1
        _ops._count_launch(self.kernels_per_replay)
        if self.steps_in_ring >= self.ring_len:
            raise RuntimeError('GraphedStep ring overflow: call drain() at least every ring_len steps')

    def drain(self):
        """One D2H copy: per-step scores since the last drain -> (COINNAverages, metrics) totals."""
        n = self.steps_in_ring
        avg, met = self.trainer.new_averages(), self.trainer.new_metrics()
        if n == 0:
            return avg, met
        rows = self.ring[:n].double().cpu()
        self.steps_in_ring = 0
        self.cursor.zero_()
        tmpl_avg, tmpl_met = self.it['averages'], self.it['metrics']
        col = 0
        counts = tmpl_avg.counts            # n per add() is static (batch shapes are)
        for (ix, _t) in tmpl_avg._pending:
            avg._values[ix] += float(rows[:, col].sum())
            col += 1
        avg._counts += counts * n
        if isinstance(tmpl_met, Prf1a) and tmpl_met._dev is not None:
            avg_counts = rows[:, col:col + 4].sum(0).round().long().numpy()
            met._host += avg_counts
        elif isinstance(tmpl_met, ConfusionMatrix) and tmpl_met._dev is not None:
            C = tmpl_met.num_classes
            met.matrix += rows[:, col:col + C * C].sum(0).view(C, C).float()
        return avg, met
