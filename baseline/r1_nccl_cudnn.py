"""R1 arm - "the reference's model with a competent library build": cuDNN/cuBLAS bf16 + NCCL + fused Adam.

BASELINE.md §3 / SURVEY §6.3 name this as the bar the hand-written data plane has to beat: the same
``nn.Module`` the reference arm trains (``ref_models.RefVBMNet`` / ``RefFSNet``), bf16 autocast (cuDNN
convolutions and BatchNorm, cuBLAS GEMMs), gradients living in ONE flat buffer that is all-reduced with
``torch.distributed.all_reduce(AVG)`` over NCCL, ``torch.optim.Adam(fused=True, capturable=True)``, the
whole step captured in ONE CUDA graph (NCCL collectives capture fine), one process per GPU.  None of this
repo's kernels run here; only the host-side staging helpers (pinned host -> device prefetch, lagged loss
read-back) are shared with our arm so both arms move the same bytes the same way.

The memory format ("its best layout") is chosen by timing: contiguous NCDHW vs channels_last_3d, three
eager steps each, faster one wins and is reported in the config.
"""
import os
import sys

import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(0, HERE)


class R1Step:
    def __init__(self, model_name, shape, batch, device, lr=1e-3, seed=11, layout='auto', graph=True):
        from ref_models import RefFSNet, RefVBMNet
        self.device, self.batch, self.shape = device, batch, tuple(shape)
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        torch.manual_seed(seed)                                  # identical replicas
        if model_name == 'vbm':
            self.model = RefVBMNet(in_ch=shape[0], input_shape=shape[1:]).to(device)
        else:
            self.model = RefFSNet(in_size=shape[0]).to(device)
        self.is3d = model_name == 'vbm'
        self.layout = 'contiguous'
        if self.is3d and layout in ('auto', 'channels_last_3d'):
            self.layout = self._pick_layout() if layout == 'auto' else 'channels_last_3d'
            if self.layout == 'channels_last_3d':
                self.model = self.model.to(memory_format=torch.channels_last_3d)
        self.model.train()
        # one flat gradient buffer: p.grad are views, the all-reduce is a single NCCL call
        params = [p for p in self.model.parameters()]
        self.flat = torch.zeros(sum(p.numel() for p in params), dtype=torch.float32, device=device)
        off = 0
        for p in params:          # same strides as the parameter (channels_last_3d weights are permuted-dense): fused Adam
            p.grad = torch.as_strided(self.flat, p.shape, p.stride(), off)   # wants identical layouts
            off += p.numel()
        self.params = params
        self.opt = torch.optim.Adam(params, lr=lr, fused=True, capturable=True)
        self.static_x = torch.zeros((batch, *self.shape), dtype=torch.float32, device=device)
        self.static_y = torch.zeros((batch,), dtype=torch.int64, device=device)
        if self.is3d and self.layout == 'channels_last_3d':
            self.static_x = self.static_x.contiguous(memory_format=torch.channels_last_3d)
        self.loss = None
        self.graph = None
        self.kernels_per_step = None
        if graph:
            self._capture()

    # -------------------------------------------------------------------------------------------
    def _fwd_bwd(self):
        with torch.autocast('cuda', dtype=torch.bfloat16):
            logits = self.model(self.static_x)
        loss = torch.nn.functional.cross_entropy(logits.float(), self.static_y)
        loss.backward()
        return loss

    def _eager_step(self):
        loss = self._fwd_bwd()
        if self.world > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG)
        self.opt.step()
        self.flat.zero_()          # grads stay bound to the flat buffer (no set_to_none)
        return loss.detach()

    def _pick_layout(self):
        best, best_ms = 'contiguous', None
        state = {k: v.clone() for k, v in self.model.state_dict().items()}
        for name in ('contiguous', 'channels_last_3d'):
            m = self.model.to(memory_format=torch.channels_last_3d) if name == 'channels_last_3d' else self.model
            x = torch.randn((self.batch, *self.shape), device=self.device)
            if name == 'channels_last_3d':
                x = x.contiguous(memory_format=torch.channels_last_3d)
            y = torch.randint(0, 2, (self.batch,), device=self.device)
            ms = []
            for i in range(4):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                with torch.autocast('cuda', dtype=torch.bfloat16):
                    out = m(x)
                torch.nn.functional.cross_entropy(out.float(), y).backward()
                e1.record()
                torch.cuda.synchronize()
                ms.append(e0.elapsed_time(e1))
                for p in m.parameters():
                    p.grad = None
            t = min(ms[1:])
            if self.world > 1:       # every rank must take the same decision
                tt = torch.tensor([t], device=self.device)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                t = float(tt)
            if best_ms is None or t < best_ms:
                best, best_ms = name, t
            self.model = self.model.to(memory_format=torch.contiguous_format)
        self.model.load_state_dict(state)
        self.layout_ms = best_ms
        return best

    def _capture(self):
        side = torch.cuda.Stream(self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        snap_p = [p.detach().clone() for p in self.params]
        snap_b = [(b, b.clone()) for b in self.model.buffers()]
        with torch.cuda.stream(side):
            for _ in range(3):                       # cuDNN autotune, NCCL channel setup, Adam state init
                self._eager_step()
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = self._eager_step()
        with torch.no_grad():                        # warm-up + capture were real updates: roll the weights back
            for p, s in zip(self.params, snap_p):
                p.copy_(s)
            for b, s in snap_b:
                b.copy_(s)
            for st in self.opt.state.values():
                for k, v in st.items():
                    if torch.is_tensor(v):
                        v.zero_()
            self.flat.zero_()

    # -------------------------------------------------------------------------------------------
    def step(self, x, y):
        self.static_x.copy_(x, non_blocking=True)
        self.static_y.copy_(y, non_blocking=True)
        if self.graph is not None:
            self.graph.replay()
        else:
            self.loss = self._eager_step()
        return self.loss


def launch_list(step, n=1):
    """Kernel names of one step (torch profiler, run OUTSIDE any timed region) - shows that the arm is the library
    path: cudnn / cublas(Lt) / nccl / ATen fused adam."""
    from torch.profiler import ProfilerActivity, profile
    x = torch.randn((step.batch, *step.shape), device=step.device)
    y = torch.randint(0, 2, (step.batch,), device=step.device)
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(n):
            step.step(x, y)
        torch.cuda.synchronize()
    rows = [(e.key, e.count, e.device_time_total) for e in prof.key_averages() if e.device_time_total > 0]
    rows.sort(key=lambda r: -r[2])
    return rows
