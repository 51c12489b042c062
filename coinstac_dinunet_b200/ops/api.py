"""Python entry points of the native ops (thin argument checking + dispatch).

Each function documents the kernel it launches (``csrc/*.cu``) and the PyTorch oracle the unit
tests compare against (``tests/test_ops_gpu.py``).
"""
import torch as _torch

__all__ = ['count_binary', 'count_confusion', 'orthogonalize_', 'softmax_nll', 'SoftmaxNLL']


def _ext():
    from . import extension
    return extension()


def count_binary(pred, true, counter):
    """counter[4] (int64: tn, fp, fn, tp) += histogram of ``2*true + pred`` (255 -> 1).
    Kernel: ``metrics.cu::count_binary_kernel`` - one pass, warp-aggregated atomics, no sync."""
    _ext().count_binary(pred.contiguous(), true.contiguous(), counter)
    return counter


def count_confusion(pred, true, matrix):
    """matrix[C, C] (int64) [pred, true] += 1.  Kernel: ``metrics.cu::count_confusion_kernel``
    (shared-memory histogram per CTA, one global atomic per non-zero bin)."""
    _ext().count_confusion(pred.contiguous(), true.contiguous(), matrix)
    return matrix


def orthogonalize_(matrix, epsilon=1e-8):
    """In-place column Gram-Schmidt of a tall [m, r] fp32 matrix, one CTA per matrix
    (``powersgd.cu::orthogonalize_kernel``)."""
    _ext().orthogonalize(matrix, float(epsilon))
    return matrix


class SoftmaxNLL(_torch.autograd.Function):
    """Fused log-softmax + NLL(mean) + argmax.  Forward returns (loss, pred); backward writes
    ``(softmax - onehot) / N`` in one pass (``loss.cu``)."""

    @staticmethod
    def forward(ctx, logits, labels):
        loss, pred, probs = _ext().softmax_nll_fwd(logits.contiguous(), labels.contiguous())
        ctx.save_for_backward(probs, labels)
        ctx.mark_non_differentiable(pred)
        return loss, pred

    @staticmethod
    def backward(ctx, g_loss, _g_pred):
        probs, labels = ctx.saved_tensors
        return _ext().softmax_nll_bwd(probs, labels, g_loss.contiguous()), None


def softmax_nll(logits, labels):
    """(mean NLL of log-softmax(logits), argmax) - native on CUDA, PyTorch elsewhere."""
    from . import native_available
    if logits.is_cuda and native_available():
        return SoftmaxNLL.apply(logits, labels)
    logp = _torch.log_softmax(logits.float(), dim=1)
    return _torch.nn.functional.nll_loss(logp, labels), logp.argmax(1)
