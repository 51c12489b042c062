"""Tracing / profiling helpers (the reference advertises per-site profiling but implements none).

* ``nvtx_range(name)``      - NVTX range when CUDA is present (shows up in Nsight), no-op elsewhere
* ``DeviceTimer``           - CUDA-event timers on the launching stream, no host sync until ``report()``
* ``site_profile(cache)``   - what ``compspec`` "profile": true maps to: per-phase device times appended to
                              ``cache['profile']`` and dumped with ``logs.json``
All multi-GPU numbers are reduced as max over ranks by the caller (bench.py), never by wall clock.
"""
import contextlib
import time

import torch


@contextlib.contextmanager
def nvtx_range(name):
    on = torch.cuda.is_available()
    if on:
        torch.cuda.nvtx.range_push(name)
    try:
        yield
    finally:
        if on:
            torch.cuda.nvtx.range_pop()


class DeviceTimer:
    """``with timer('fwd'): ...`` records a CUDA event pair per use; ``report()`` synchronises once."""

    def __init__(self, device=None):
        self.cuda = torch.cuda.is_available()
        self.device = device
        self._open, self._pairs, self._host = {}, {}, {}

    @contextlib.contextmanager
    def __call__(self, name):
        if self.cuda:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            with nvtx_range(name):
                yield
            e1.record()
            self._pairs.setdefault(name, []).append((e0, e1))
        else:
            t0 = time.perf_counter()
            yield
            self._host.setdefault(name, []).append((time.perf_counter() - t0) * 1e3)

    def report(self, reset=True):
        """{name: {'count', 'total_ms', 'mean_ms'}}"""
        out = {}
        if self.cuda and self._pairs:
            torch.cuda.synchronize(self.device)
        for name, pairs in self._pairs.items():
            ms = [a.elapsed_time(b) for a, b in pairs]
            out[name] = {'count': len(ms), 'total_ms': sum(ms), 'mean_ms': sum(ms) / len(ms)}
        for name, ms in self._host.items():
            out[name] = {'count': len(ms), 'total_ms': sum(ms), 'mean_ms': sum(ms) / len(ms)}
        if reset:
            self._pairs, self._host = {}, {}
        return out


def site_profile(cache, timer, key='profile'):
    """Append the timer report to ``cache[key]`` (JSON-able; lands in ``logs.json``)."""
    rep = timer.report()
    if rep:
        cache.setdefault(key, []).append(rep)
    return rep
