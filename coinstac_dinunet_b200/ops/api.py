"""Python entry points of the native ops (argument checking + dispatch to the C ABI).

Each function names the kernel it launches (``csrc/*.cu``) and has a PyTorch oracle that the GPU
tests compare against (``tests/test_ops_gpu.py``).
"""
import torch as _torch

__all__ = ['count_binary', 'count_confusion', 'orthogonalize_', 'softmax_nll', 'SoftmaxNLL']


def _native():
    from . import native
    return native


def _bump(n=1):
    from . import _count_launch
    _count_launch(n)


def _int_code(t):
    nat = _native()
    if t.dtype not in nat.INT_CODES:
        t = t.long()
    return t.contiguous(), nat.INT_CODES[t.dtype]


def _count(pred, true, out, C, confusion):
    nat = _native()
    assert out.dtype == _torch.int64 and out.is_cuda and out.is_contiguous()
    pred, pc = _int_code(pred.reshape(-1))
    true, tc = _int_code(true.reshape(-1))
    assert pred.numel() == true.numel()
    nat.check(nat.lib().coinn_count(pred.data_ptr(), pc, true.data_ptr(), tc, out.data_ptr(), pred.numel(),
                                    C, int(confusion), nat.stream_ptr(pred.device)), 'coinn_count')
    _bump()
    return out


def count_binary(pred, true, counter):
    """counter[4] (int64: tn, fp, fn, tp) += histogram of ``2*true + pred`` (255 -> 1).
    Kernel ``metrics.cu::count_binary_kernel``: one pass, warp-aggregated atomics, no host sync."""
    return _count(pred, true, counter, 2, False)


def count_confusion(pred, true, matrix):
    """matrix[C, C] (int64) ``[pred, true] += 1``.  Kernel ``metrics.cu::count_confusion_kernel``
    (shared-memory histogram per CTA, one global atomic per non-zero bin)."""
    return _count(pred, true, matrix, matrix.shape[0], True)


def orthogonalize_(matrix, epsilon=1e-8):
    """In-place column Gram-Schmidt of a tall [m, r<=32] fp32 matrix in one launch
    (``powersgd.cu::orthogonalize_kernel``)."""
    nat = _native()
    assert matrix.is_cuda and matrix.dtype == _torch.float32 and matrix.dim() == 2 and matrix.is_contiguous()
    nat.check(nat.lib().coinn_orthogonalize(matrix.data_ptr(), matrix.shape[0], matrix.shape[1], float(epsilon),
                                            nat.stream_ptr(matrix.device)), 'coinn_orthogonalize')
    _bump()
    return matrix


class SoftmaxNLL(_torch.autograd.Function):
    """Fused log-softmax + NLL(mean) + argmax (``loss.cu``).  forward -> (loss, pred); backward writes
    ``(softmax - onehot) * g / N`` in one pass, in the dtype of the logits."""

    @staticmethod
    def forward(ctx, logits, labels):
        nat = _native()
        logits = logits.contiguous()
        labels = labels.contiguous().long()
        n, c = logits.shape
        probs = _torch.empty((n, c), dtype=_torch.float32, device=logits.device)
        pred = _torch.empty((n,), dtype=_torch.int64, device=logits.device)
        loss = _torch.zeros((), dtype=_torch.float32, device=logits.device)
        nat.check(nat.lib().coinn_softmax_nll_fwd(logits.data_ptr(), labels.data_ptr(), probs.data_ptr(),
                                                  pred.data_ptr(), loss.data_ptr(), n, c,
                                                  nat.FLOAT_CODES[logits.dtype], nat.stream_ptr(logits.device)),
                  'coinn_softmax_nll_fwd')
        _bump()
        ctx.save_for_backward(probs, labels)
        ctx.in_dtype = logits.dtype
        ctx.mark_non_differentiable(pred)
        return loss, pred

    @staticmethod
    def backward(ctx, g_loss, _g_pred):
        nat = _native()
        probs, labels = ctx.saved_tensors
        n, c = probs.shape
        g = g_loss.contiguous().float().reshape(1)
        dlogits = _torch.empty((n, c), dtype=ctx.in_dtype, device=probs.device)
        nat.check(nat.lib().coinn_softmax_nll_bwd(probs.data_ptr(), labels.data_ptr(), g.data_ptr(),
                                                  dlogits.data_ptr(), n, c, nat.FLOAT_CODES[ctx.in_dtype],
                                                  nat.stream_ptr(probs.device)), 'coinn_softmax_nll_bwd')
        _bump()
        return dlogits, None


def softmax_nll(logits, labels):
    """``(mean NLL of log_softmax(logits), argmax)`` - native on CUDA, PyTorch oracle elsewhere."""
    from . import native_available
    if logits.is_cuda and logits.dim() == 2 and logits.dtype in _native().FLOAT_CODES and native_available():
        return SoftmaxNLL.apply(logits, labels)
    logp = _torch.log_softmax(logits.float(), dim=1)
    return _torch.nn.functional.nll_loss(logp, labels), logp.argmax(1)
