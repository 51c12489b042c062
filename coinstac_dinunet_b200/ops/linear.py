"""Linear layers on the hand-written tcgen05 GEMM (``csrc/gemm_tcgen05.cu``).

``gemm_tn(A, B)`` computes ``A[M,K] @ B[N,K].T`` in bf16 with fp32 accumulation in TMEM.  The three
GEMMs of a Linear layer map onto it as
    forward  y  = gemm_tn(x,  W)                      [M,N]   (+bias, +ReLU fused in the epilogue)
    dgrad    dx = gemm_tn(dy, W^T contiguous)         [M,K]
    wgrad    dW = gemm_tn(dy^T, x^T) in fp32          [N,K]
Small-M problems (batch 8-16) are split along K so that ~148 CTAs stream the weight matrix.
"""
import torch as _torch

from . import native as _nat

_SMS = 148


def _bump():
    from . import _count_launch
    _count_launch()


def gemm_tn(a, b, bias=None, relu=False, out_dtype=_torch.bfloat16, bias_mode=1, split_k=None, out=None):
    """``a[M,K] @ b[N,K].T`` -> ``[M,N]``.  a, b: bf16, row-major, K % 8 == 0."""
    assert a.is_cuda and b.is_cuda and a.dtype == _torch.bfloat16 and b.dtype == _torch.bfloat16
    M, K = a.shape
    N, K2 = b.shape
    assert K == K2, (a.shape, b.shape)
    if K % 8:  # TMA needs 16-byte aligned row strides: pad K with zeros (exact)
        pad = 8 - K % 8
        a = _torch.nn.functional.pad(a, (0, pad))
        b = _torch.nn.functional.pad(b, (0, pad))
        K += pad
    a, b = a.contiguous(), b.contiguous()
    if split_k is None:
        tiles = ((M + 127) // 128) * ((N + 127) // 128)
        kblocks = (K + 63) // 64
        split_k = max(1, min(kblocks // 4, _SMS // tiles)) if tiles * 2 <= _SMS else 1
    fused_epilogue = split_k <= 1
    if fused_epilogue:
        c = out if out is not None else _torch.empty((M, N), dtype=out_dtype, device=a.device)
        code = 0 if c.dtype == _torch.bfloat16 else 1
        assert c.dtype in (_torch.bfloat16, _torch.float32) and c.is_contiguous()
        bias_t = bias.float().contiguous() if bias is not None else None
        _nat.check(_nat.lib().coinn_gemm_bf16_tn(a.data_ptr(), b.data_ptr(), c.data_ptr(),
                                                 bias_t.data_ptr() if bias_t is not None else None,
                                                 M, N, K, K, K, N, code, int(relu), int(bias_mode), 1,
                                                 _nat.stream_ptr(a.device)), 'coinn_gemm_bf16_tn')
        _bump()
        return c
    acc = _torch.zeros((M, N), dtype=_torch.float32, device=a.device)
    _nat.check(_nat.lib().coinn_gemm_bf16_tn(a.data_ptr(), b.data_ptr(), acc.data_ptr(), None, M, N, K, K, K, N,
                                             1, 0, 0, int(split_k), _nat.stream_ptr(a.device)), 'coinn_gemm_bf16_tn')
    _bump()
    if bias is not None:
        acc = acc + (bias.float() if bias_mode == 1 else bias.float().unsqueeze(1))
    if relu:
        acc = acc.relu_()
    res = acc if out_dtype == _torch.float32 else acc.to(out_dtype)
    if out is not None:
        out.copy_(res)
        return out
    return res


class LinearFn(_torch.autograd.Function):
    """y = relu?(x @ W^T + b) with all three GEMMs on tcgen05.  x: [M,K] any float dtype."""

    @staticmethod
    def forward(ctx, x, weight, bias, relu):
        xb = x.to(_torch.bfloat16).contiguous()
        wb = weight.to(_torch.bfloat16).contiguous()
        y = gemm_tn(xb, wb, bias=bias, relu=relu, out_dtype=_torch.bfloat16)
        ctx.save_for_backward(xb, wb, y if relu else None)
        ctx.relu, ctx.has_bias = relu, bias is not None
        ctx.in_dtype, ctx.w_dtype = x.dtype, weight.dtype
        return y

    @staticmethod
    def backward(ctx, dy):
        xb, wb, y = ctx.saved_tensors
        dy = dy.to(_torch.bfloat16)
        if ctx.relu:
            dy = dy * (y > 0)
        dy = dy.contiguous()
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = gemm_tn(dy, wb.t().contiguous(), out_dtype=_torch.bfloat16).to(ctx.in_dtype)
        if ctx.needs_input_grad[1]:
            dw = gemm_tn(dy.t().contiguous(), xb.t().contiguous(), out_dtype=_torch.float32).to(ctx.w_dtype)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = dy.float().sum(0)
        return dx, dw, db, None


DIRECT_GRAD_DISABLED = False      # debugging switch: force every gradient through autograd's AccumulateGrad

# Kernels that accumulate straight into ``param.grad`` hand ``None`` to autograd, so ``post_accumulate_grad`` hooks never
# fire for those parameters.  Whoever needs to know "this gradient is final" (DistArena's bucketed backward overlap)
# registers a listener here and the direct-mode kernels call ``notify_grad_written`` right after their launch.
import weakref as _weakref

_GRAD_LISTENERS = {}      # id(param) -> (weakref(param), callable); tensors cannot key a Weak*Dictionary (== is elementwise)


def register_grad_listener(param, fn):
    key = id(param)
    _GRAD_LISTENERS[key] = (_weakref.ref(param, lambda _r, k=key: _GRAD_LISTENERS.pop(k, None)), fn)


def notify_grad_written(*params):
    if not _GRAD_LISTENERS:
        return
    for q in params:
        if q is None:
            continue
        hit = _GRAD_LISTENERS.get(id(q))
        if hit is not None and hit[0]() is q:
            hit[1]()


def direct_grad_ok(param):
    """True when a backward kernel may accumulate straight into ``param.grad`` (fp32, contiguous, already allocated -
    e.g. a view of the DistArena gradient arena) and return ``None`` to autograd: no temporary, no AccumulateGrad add.
    COINN_DIRECT_GRAD=0 disables it (needed if post-accumulate-grad hooks must fire for every parameter)."""
    import os
    g = getattr(param, 'grad', None)
    return (not DIRECT_GRAD_DISABLED and os.environ.get('COINN_DIRECT_GRAD', '1') != '0' and g is not None and param.dtype == _torch.float32
            and g.dtype == _torch.float32 and g.is_contiguous() and param.is_contiguous())


SMALL_M = 32


class SmallLinearFn(_torch.autograd.Function):
    """y = relu?(x @ W^T + b) for M <= 32 rows on the CUDA-core kernels of ``csrc/linear_small.cu`` (two launches per
    layer and step instead of ~15; fp32 weights are streamed once, dW/db accumulate in place into ``.grad``)."""

    @staticmethod
    def forward(ctx, x, weight, bias, relu):
        if x.dtype not in (_torch.float32, _torch.bfloat16):
            x = x.float()
        x = x.contiguous()
        w = weight.detach()
        w = w if (w.dtype == _torch.float32 and w.is_contiguous()) else w.float().contiguous()
        b = None if bias is None else bias.detach().float().contiguous()
        M, K = x.shape
        N = w.shape[0]
        y = _torch.empty((M, N), dtype=_torch.float32, device=x.device)
        _nat.check(_nat.lib().coinn_linear_small_fwd(x.data_ptr(), 1 if x.dtype == _torch.bfloat16 else 0, w.data_ptr(),
                                                     b.data_ptr() if b is not None else None, y.data_ptr(), M, N, K, int(relu),
                                                     _nat.stream_ptr(x.device)), 'coinn_linear_small_fwd')
        _bump()
        ctx.save_for_backward(x, w, y if relu else None)
        ctx.params = (weight, bias)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, y = ctx.saved_tensors
        weight, bias = ctx.params
        M, K = x.shape
        N = w.shape[0]
        dy = dy.float().contiguous()
        need_w, need_b = ctx.needs_input_grad[1], bias is not None and ctx.needs_input_grad[2]
        direct = need_w and direct_grad_ok(weight) and w.data_ptr() == weight.data_ptr() and (not need_b or direct_grad_ok(bias))
        if direct:
            dw, db = weight.grad, (bias.grad if need_b else None)
        else:
            dw = _torch.zeros((N, K), dtype=_torch.float32, device=x.device)
            db = _torch.zeros(N, dtype=_torch.float32, device=x.device) if need_b else None
        dx = _torch.zeros((M, K), dtype=_torch.float32, device=x.device) if ctx.needs_input_grad[0] else None
        _nat.check(_nat.lib().coinn_linear_small_bwd(dy.data_ptr(), y.data_ptr() if y is not None else None, x.data_ptr(),
                                                     1 if x.dtype == _torch.bfloat16 else 0, w.data_ptr(), dw.data_ptr(),
                                                     db.data_ptr() if db is not None else None,
                                                     dx.data_ptr() if dx is not None else None, M, N, K,
                                                     _nat.stream_ptr(x.device)), 'coinn_linear_small_bwd')
        _bump()
        if dx is not None and x.dtype != _torch.float32:
            dx = dx.to(x.dtype)
        if direct:
            notify_grad_written(weight, bias if need_b else None)
            return dx, None, None, None
        return dx, dw.to(weight.dtype), (db.to(bias.dtype) if db is not None else None), None


class SmallLinearBnReluFn(_torch.autograd.Function):
    """z = relu?(BatchNorm1d(x @ W^T + b)) for M <= 32 rows in ONE launch each way (``csrc/linear_small.cu``): at these
    batch sizes the CTA that owns an output feature owns its whole batch column, so the batch statistics, the
    normalisation, the ReLU and the running-statistics update are the epilogue of the Linear kernel, and BatchNorm's
    backward is the prologue of the Linear backward.  Replaces Linear + ATen batch_norm + ReLU (3 forward / ~6 backward
    launches plus dtype casts) in the FreeSurfer MLP (ref model: README.md:31)."""

    @staticmethod
    def forward(ctx, x, weight, bias, gamma, beta, running_mean, running_var, nbt, eps, momentum, training, relu):
        if x.dtype not in (_torch.float32, _torch.bfloat16):
            x = x.float()
        x = x.contiguous()
        w = weight.detach()
        w = w if (w.dtype == _torch.float32 and w.is_contiguous()) else w.float().contiguous()
        b = None if bias is None else bias.detach().float().contiguous()
        g, be = gamma.detach().float().contiguous(), beta.detach().float().contiguous()
        M, K = x.shape
        N = w.shape[0]
        z = _torch.empty((M, N), dtype=_torch.float32, device=x.device)
        xhat = _torch.empty((M, N), dtype=_torch.float32, device=x.device) if training else None
        invstd = _torch.empty(N, dtype=_torch.float32, device=x.device) if training else None
        ptr = lambda t: t.data_ptr() if t is not None else None
        _nat.check(_nat.lib().coinn_linear_bn_small_fwd(
            x.data_ptr(), 1 if x.dtype == _torch.bfloat16 else 0, w.data_ptr(), ptr(b), g.data_ptr(), be.data_ptr(),
            ptr(running_mean), ptr(running_var), ptr(nbt) if training else None, ptr(xhat), ptr(invstd), z.data_ptr(),
            M, N, K, int(relu), int(bool(training)), float(eps), float(momentum), _nat.stream_ptr(x.device)),
            'coinn_linear_bn_small_fwd')
        _bump()
        ctx.save_for_backward(x, w, xhat, invstd, g, be)
        ctx.params = (weight, bias, gamma, beta)
        ctx.relu, ctx.training = bool(relu), bool(training)
        return z

    @staticmethod
    def backward(ctx, dz):
        if not ctx.training:
            raise RuntimeError('SmallLinearBnReluFn.backward is only defined for training-mode BatchNorm')
        x, w, xhat, invstd, g, be = ctx.saved_tensors
        weight, bias, gamma, beta = ctx.params
        M, K = x.shape
        N = w.shape[0]
        dz = dz.float().contiguous()
        has_b = bias is not None and ctx.needs_input_grad[2]
        direct = direct_grad_ok(weight) and w.data_ptr() == weight.data_ptr() and direct_grad_ok(gamma) and \
            direct_grad_ok(beta) and (not has_b or direct_grad_ok(bias))
        if direct:
            dw, db, dg, dbe = weight.grad, (bias.grad if has_b else None), gamma.grad, beta.grad
        else:
            dw = _torch.zeros((N, K), dtype=_torch.float32, device=x.device)
            db = _torch.zeros(N, dtype=_torch.float32, device=x.device) if has_b else None
            dg, dbe = _torch.zeros(N, dtype=_torch.float32, device=x.device), _torch.zeros(N, dtype=_torch.float32, device=x.device)
        dx = _torch.zeros((M, K), dtype=_torch.float32, device=x.device) if ctx.needs_input_grad[0] else None
        ptr = lambda t: t.data_ptr() if t is not None else None
        _nat.check(_nat.lib().coinn_linear_bn_small_bwd(
            dz.data_ptr(), xhat.data_ptr(), invstd.data_ptr(), g.data_ptr(), be.data_ptr(), x.data_ptr(),
            1 if x.dtype == _torch.bfloat16 else 0, w.data_ptr(), dw.data_ptr(), ptr(db), dg.data_ptr(), dbe.data_ptr(), ptr(dx),
            M, N, K, int(ctx.relu), _nat.stream_ptr(x.device)), 'coinn_linear_bn_small_bwd')
        _bump()
        if dx is not None and x.dtype != _torch.float32:
            dx = dx.to(x.dtype)
        tail = (None,) * 7
        if direct:
            notify_grad_written(weight, bias if has_b else None, gamma, beta)
            return (dx, None, None, None, None) + tail
        return (dx, dw.to(weight.dtype), (db.to(bias.dtype) if db is not None else None), dg.to(gamma.dtype),
                dbe.to(beta.dtype)) + tail


def linear_bn_relu(x, lin, bn, relu=True):
    """``relu?(bn(lin(x)))`` for an ``nn.Linear`` + ``nn.BatchNorm1d`` pair: fused small-batch kernel when it applies
    (CUDA, M <= 32, affine BatchNorm with running statistics or training mode), PyTorch composition otherwise."""
    x2 = x.reshape(-1, x.shape[-1])
    fused = (x2.is_cuda and x2.shape[0] <= SMALL_M and bn.affine and lin.weight.dtype == _torch.float32
             and (bn.training or bn.track_running_stats) and (not bn.training or x2.shape[0] > 1))
    if not fused:
        y = bn(_torch.nn.functional.linear(x2.to(lin.weight.dtype), lin.weight, lin.bias))
        return y.relu() if relu else y
    use_batch = bn.training or not bn.track_running_stats
    mom = bn.momentum if bn.momentum is not None else 0.1
    tracked = bn.track_running_stats
    return SmallLinearBnReluFn.apply(x2, lin.weight, lin.bias, bn.weight, bn.bias,
                                     bn.running_mean if tracked else None, bn.running_var if tracked else None,
                                     bn.num_batches_tracked if tracked else None, bn.eps, mom, use_batch, relu)


def linear(x, weight, bias, relu):
    """Dispatch: small-batch CUDA-core kernels (M <= 32) or the tcgen05 GEMM path."""
    if x.shape[0] <= SMALL_M:
        return SmallLinearFn.apply(x, weight, bias, relu)
    return LinearFn.apply(x, weight, bias, relu)


class B200Linear(_torch.nn.Linear):
    """Drop-in ``nn.Linear`` whose forward/backward run on the tcgen05 GEMM (optionally with the
    following ReLU fused into the epilogue)."""
    is_native = True

    def __init__(self, in_features, out_features, bias=True, fuse_relu=False, **kw):
        super().__init__(in_features, out_features, bias=bias, **kw)
        self.fuse_relu = fuse_relu

    @classmethod
    def from_linear(cls, lin, fuse_relu=False):
        new = cls.__new__(cls)
        _torch.nn.Module.__init__(new)
        new.in_features, new.out_features = lin.in_features, lin.out_features
        new.weight, new.bias = lin.weight, lin.bias
        new.fuse_relu = fuse_relu
        return new

    def forward(self, x):
        if not x.is_cuda:
            y = _torch.nn.functional.linear(x, self.weight, self.bias)
            return y.relu() if self.fuse_relu else y
        lead = x.shape[:-1]
        y = linear(x.reshape(-1, x.shape[-1]), self.weight, self.bias, self.fuse_relu)
        return y.reshape(*lead, self.out_features)
