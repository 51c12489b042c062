"""Serialisable metrics with device-resident accumulation.

API parity with coinstac_dinunet/metrics/metrics.py:17-329 (``COINNMetrics`` interface,
``COINNAverages``, ``Prf1a``, ``ConfusionMatrix``, ``AUCROCMetrics``) including the wire
format of ``serialize()`` / ``reduce_sites()``.

What is different (B200-first):

* ``add()`` never forces a host sync.  CUDA inputs are folded into a small device
  counter tensor by one fused kernel (``ops.count_binary`` / ``ops.count_confusion``;
  a ``torch.bincount`` fallback is used on CPU).  The reference pays four ``.item()``
  syncs per ``Prf1a.add`` (metrics.py:158-170) and one per loss (README:82).
* Host numbers are materialised lazily - the first time a score is *read*.
* ``ConfusionMatrix`` aggregation on the remote works (reference quirk §8.5-7).
"""
import abc as _abc
import time as _time
import typing as _typing

import numpy as _np
import torch as _torch

from ..config import metrics_eps as _eps, metrics_num_precision as _nump


def _ops():
    """Native kernels if the extension is loaded *and* usable, else None."""
    try:
        from .. import ops as _o
        return _o if _o.native_available() else None
    except Exception:
        return None


class COINNMetrics:
    """Interface every metric implements (ref metrics.py:17-84)."""

    def __init__(self, device='cpu', **kw):
        self.device = device

    @_abc.abstractmethod
    def add(self, *args, **kw):
        raise NotImplementedError('Must be implemented.')

    def accumulate(self, other):
        pass

    def reset(self):
        pass

    def get(self, *args, **kw) -> _typing.List[float]:
        return [0.0]

    @property
    def eps(self):
        return _eps

    @property
    def num_precision(self):
        return _nump

    @property
    def time(self):
        return _time.time()

    def extract(self, name):
        """Look a score up by (case-insensitive) name; call it if it is a method."""
        attr = getattr(self, str(name).lower())
        return attr() if callable(attr) else attr

    @_abc.abstractmethod
    def serialize(self, **kw):
        pass

    @_abc.abstractmethod
    def reduce_sites(self, scores):
        pass


class COINNAverages(COINNMetrics):
    """K running weighted means (losses).  ``serialize() == [values, counts]``.

    ``add(val, n, index)`` accepts Python numbers *or* 0-d tensors.  Tensors are kept
    on their device (``value * n`` is appended to a pending list) and reduced with a
    single D2H copy when a number is first needed - so a training step can call
    ``averages.add(loss.detach(), n)`` without stalling the stream.
    """

    def __init__(self, num_averages=1, **kw):
        super().__init__(**kw)
        self.num_averages = int(num_averages)
        self._values = _np.zeros(self.num_averages, dtype=_np.float64)
        self._counts = _np.zeros(self.num_averages, dtype=_np.float64)
        self._pending = []  # (index, tensor_of_val_times_n)

    # --- lazy device -> host -------------------------------------------------
    def _flush(self):
        if not self._pending:
            return
        pend, self._pending = self._pending, []
        stacked = _torch.stack([t.reshape(()).double() for _, t in pend]).cpu().numpy()
        for (ix, _), v in zip(pend, stacked):
            self._values[ix] += float(v)

    @property
    def values(self):
        self._flush()
        return self._values

    @values.setter
    def values(self, v):
        self._pending = []
        self._values = _np.asarray(v, dtype=_np.float64)

    @property
    def counts(self):
        return self._counts

    @counts.setter
    def counts(self, v):
        self._counts = _np.asarray(v, dtype=_np.float64)

    # --- API -------------------------------------------------------------------
    def add(self, val=0, n=1, index=0):
        if isinstance(val, _torch.Tensor):
            if val.device.type == 'cpu':
                self._values[index] += float(val) * n
            else:
                self._pending.append((index, val.detach() * n))
        else:
            self._values[index] += val * n
        self._counts[index] += n

    def accumulate(self, other):
        self._values += other._values
        self._counts += other._counts
        self._pending.extend(other._pending)

    def reset(self):
        self._values = _np.zeros(self.num_averages, dtype=_np.float64)
        self._counts = _np.zeros(self.num_averages, dtype=_np.float64)
        self._pending = []

    def get(self):
        denom = _np.where(self._counts == 0, _np.inf, self._counts)
        return _np.round(self.values / denom, self.num_precision)

    def average(self, reduce_mean=True):
        avgs = self.get()
        if reduce_mean:
            return round(float(sum(avgs)) / len(avgs), self.num_precision)
        return avgs

    def serialize(self, **kw):
        return [self.values.tolist(), self._counts.tolist()]

    def reduce_sites(self, scores: list):
        """Exact weighted mean across sites: element-wise *sum* of [values, counts]."""
        if len(scores) == 0:
            return
        total = _np.asarray(scores, dtype=_np.float64).sum(0)
        self.values, self.counts = total[0], total[1]


class Prf1a(COINNMetrics):
    """Binary precision / recall / F1 / accuracy / IoU from TP, FP, TN, FN.

    Counting uses the ``2*true + pred`` coding of the reference (metrics.py:158-170),
    with 255 treated as 1 so 8-bit masks work.  Counts live in a 4-element device
    tensor ``[tn, fp, fn, tp]`` (the natural bincount order of that coding) until read.
    """

    def __init__(self, **kw):
        super().__init__(**kw)
        self._host = _np.zeros(4, dtype=_np.int64)  # tn, fp, fn, tp
        self._dev = None
        self._precision = 0
        self._recall = 0
        self._accuracy = 0

    # --- counts ----------------------------------------------------------------
    def _flush(self):
        if self._dev is not None:
            self._host += self._dev.cpu().numpy().astype(_np.int64)
            self._dev = None

    def _count(self, which):
        self._flush()
        return int(self._host[which])

    def _set(self, which, v):
        self._flush()
        self._host[which] = int(v)

    tn = property(lambda s: s._count(0), lambda s, v: s._set(0, v))
    fp = property(lambda s: s._count(1), lambda s, v: s._set(1, v))
    fn = property(lambda s: s._count(2), lambda s, v: s._set(2, v))
    tp = property(lambda s: s._count(3), lambda s, v: s._set(3, v))

    def add(self, pred, true):
        pred = pred.detach().reshape(-1)
        true = true.detach().reshape(-1)
        if pred.is_cuda:
            ops = _ops()
            if self._dev is None or self._dev.device != pred.device:
                self._flush()
                self._dev = _torch.zeros(4, dtype=_torch.int64, device=pred.device)
            if ops is not None:
                ops.count_binary(pred, true, self._dev)  # one kernel, no sync
                return
            codes = self._codes(pred, true)
            self._dev += _torch.bincount(codes, minlength=4)[:4]
        else:
            codes = self._codes(pred, true)
            self._host += _torch.bincount(codes, minlength=4)[:4].numpy()

    @staticmethod
    def _codes(pred, true):
        p = pred.to(_torch.int64)
        t = true.to(_torch.int64)
        p = _torch.where(p == 255, _torch.ones_like(p), p)
        t = _torch.where(t == 255, _torch.ones_like(t), t)
        codes = 2 * t + p
        return codes[(codes >= 0) & (codes <= 3)]

    def accumulate(self, other):
        other._flush()
        if other._dev is not None:  # pragma: no cover
            other._flush()
        self._flush()
        self._host += other._host

    def reset(self):
        self._host = _np.zeros(4, dtype=_np.int64)
        self._dev = None

    # --- scores ----------------------------------------------------------------
    def _r(self, x):
        # NOT round(float(x)): after reduce_sites the reduced score is a NumPy scalar and the reference rounds it with
        # NumPy's rule (metrics.py:183-195) - the two rules differ on half-way cases such as 0.484105 -> 0.4841 / 0.48411
        return round(x, self.num_precision)

    @property
    def precision(self):
        p = self.tp / max(self.tp + self.fp, self.eps)
        return self._r(max(p, self._precision))

    @property
    def recall(self):
        r = self.tp / max(self.tp + self.fn, self.eps)
        return self._r(max(r, self._recall))

    @property
    def accuracy(self):
        a = (self.tp + self.tn) / max(self.tp + self.fp + self.fn + self.tn, self.eps)
        return self._r(max(a, self._accuracy))

    @property
    def f1(self):
        return self.f_beta(beta=1)

    def f_beta(self, beta=1):
        p, r, b2 = self.precision, self.recall, beta ** 2
        return self._r((1 + b2) * p * r / max(b2 * p + r, self.eps))

    @property
    def overlap(self):
        return self._r(self.tp / max(self.tp + self.fp + self.fn, self.eps))

    def get(self):
        return [self.accuracy, self.f1, self.precision, self.recall]

    def serialize(self, **kw):
        return [self.accuracy, self.precision, self.recall]

    def reduce_sites(self, scores: list):
        """Unweighted mean of per-site ``[accuracy, precision, recall]`` (ref metrics.py:217-218)."""
        if len(scores) == 0:
            return
        self._accuracy, self._precision, self._recall = _np.asarray(scores, dtype=_np.float64).mean(0)


class ConfusionMatrix(COINNMetrics):
    """Multi-class confusion matrix; ``matrix[pred, true]`` convention of the reference.

    ``precision``/``recall``/``f1`` are macro averages.  After ``reduce_sites`` the object
    carries site-averaged scalars which are surfaced through ``max(local, reduced)``
    exactly like ``Prf1a`` (this is the aggregation the reference intended but breaks on,
    SURVEY §8.5-7).
    """

    def __init__(self, num_classes=None, device='cpu', **kw):
        super().__init__(device, **kw)
        self.num_classes = int(num_classes)
        self.matrix = _torch.zeros(self.num_classes, self.num_classes, dtype=_torch.float32)
        self._dev = None
        self._precision = 0.0
        self._recall = 0.0
        self._accuracy = 0.0

    def _flush(self):
        if self._dev is not None:
            self.matrix += self._dev.to('cpu', _torch.float32)
            self._dev = None

    def reset(self):
        self.matrix = _torch.zeros(self.num_classes, self.num_classes, dtype=_torch.float32)
        self._dev = None

    def accumulate(self, other):
        other._flush()
        self._flush()
        self.matrix += other.matrix

    def add(self, pred: _torch.Tensor, true: _torch.Tensor):
        C = self.num_classes
        pred = pred.detach().reshape(-1)
        true = true.detach().reshape(-1)
        if pred.is_cuda:
            if self._dev is None or self._dev.device != pred.device:
                self._flush()
                self._dev = _torch.zeros(C * C, dtype=_torch.int64, device=pred.device).view(C, C)
            ops = _ops()
            if ops is not None:
                ops.count_confusion(pred, true, self._dev)
            else:
                flat = pred.long() * C + true.long()
                self._dev += _torch.bincount(flat, minlength=C * C)[:C * C].view(C, C)
        else:
            flat = pred.long() * C + true.long()
            self.matrix += _torch.bincount(flat, minlength=C * C)[:C * C].view(C, C).float()

    # --- scores ----------------------------------------------------------------
    def _per_class(self, axis_sum):
        self._flush()
        diag = self.matrix.diag().double()
        denom = _torch.clamp(axis_sum.double(), min=self.eps)
        return (diag / denom).tolist()

    def precision(self, average=True):
        self._flush()
        per = self._per_class(self.matrix.sum(0))  # column i: everything whose 2nd index is i
        if not average:
            return per
        return max(sum(per) / self.num_classes, float(self._precision))

    def recall(self, average=True):
        self._flush()
        per = self._per_class(self.matrix.sum(1))
        if not average:
            return per
        return max(sum(per) / self.num_classes, float(self._recall))

    def f1(self, average=True):
        ps = [self.precision(True)] if average else self.precision(False)
        rs = [self.recall(True)] if average else self.recall(False)
        f = _np.array([2 * p * r / max(p + r, self.eps) for p, r in zip(ps, rs)])
        return float(f[0]) if average else f

    def accuracy(self):
        self._flush()
        tot = max(float(self.matrix.sum()), self.eps)
        return max(float(self.matrix.trace()) / tot, float(self._accuracy))

    def get(self):
        r = lambda x: round(float(x), self.num_precision)
        return [r(self.accuracy()), r(self.f1()), r(self.precision()), r(self.recall())]

    def serialize(self, **kw):
        return [float(self.accuracy()), float(self.precision()), float(self.recall())]

    def reduce_sites(self, scores: list):
        if len(scores) == 0:
            return
        acc, prec, rec = _np.asarray(scores, dtype=_np.float64).mean(0)
        self._accuracy, self._precision, self._recall = float(acc), float(prec), float(rec)


class AUCROCMetrics(COINNMetrics):
    """Binary ROC-AUC.  Scores and labels stay on the device in chunk lists; the AUC is a
    sort + trapezoid on the device (ties handled like ``sklearn.metrics.roc_curve``),
    computed once when read.  The reference moves every batch to host lists
    (metrics.py:321-323) and calls sklearn (312-316)."""

    def __init__(self, device='cpu', **kw):
        super().__init__(device, **kw)
        self._prob_chunks = []
        self._label_chunks = []
        self.fpr = self.tpr = self.thresholds = None
        self._auc = 0

    # ``probabilities`` / ``labels`` are exposed as host lists for API compatibility.
    @property
    def probabilities(self):
        return _torch.cat(self._prob_chunks).cpu().tolist() if self._prob_chunks else []

    @property
    def labels(self):
        return _torch.cat(self._label_chunks).cpu().tolist() if self._label_chunks else []

    def accumulate(self, other):
        self._prob_chunks += other._prob_chunks
        self._label_chunks += other._label_chunks

    def reset(self):
        self._prob_chunks, self._label_chunks = [], []

    def add(self, pred: _torch.Tensor, true: _torch.Tensor):
        self._prob_chunks.append(pred.detach().reshape(-1).float())
        self._label_chunks.append(true.detach().reshape(-1).long())

    @staticmethod
    def _auc_from(scores, labels):
        """Area under the ROC curve; distinct-threshold trapezoid (== sklearn)."""
        dev = scores.device
        labels = labels.to(dev)
        order = _torch.argsort(scores, descending=True, stable=True)
        s, y = scores[order], (labels[order] == 1).double()
        tps, fps = _torch.cumsum(y, 0), _torch.cumsum(1 - y, 0)
        last = _torch.ones_like(s, dtype=_torch.bool)
        last[:-1] = s[1:] != s[:-1]  # keep the last element of every run of ties
        tps, fps = tps[last], fps[last]
        zero = _torch.zeros(1, dtype=_torch.double, device=dev)
        tps, fps = _torch.cat([zero, tps]), _torch.cat([zero, fps])
        P, N = tps[-1], fps[-1]
        if P <= 0 or N <= 0:
            return float('nan'), None, None
        tpr, fpr = tps / P, fps / N
        return float(_torch.trapezoid(tpr, fpr)), fpr, tpr

    def auc(self):
        if self._auc <= 0 and self._label_chunks:
            dev = self._prob_chunks[0].device
            scores = _torch.cat([c.to(dev) for c in self._prob_chunks])
            labels = _torch.cat([c.to(dev) for c in self._label_chunks])
            val, self.fpr, self.tpr = self._auc_from(scores, labels)
            return val
        return self._auc

    def get(self, *args, **kw):
        return [round(self.auc(), self.num_precision)]

    def serialize(self, **kw):
        return [self.auc()]

    def reduce_sites(self, scores: list):
        if len(scores) == 0:
            return
        self._auc = float(_np.asarray(scores, dtype=_np.float64).mean())
