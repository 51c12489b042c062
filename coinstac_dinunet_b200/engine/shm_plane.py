"""Shared-memory mailbox for the JSON control plane of a single-node ``DistEngine``.

Per engine round every site sends one small dict to the aggregator and gets one back.  Over ``torch.distributed`` object
collectives that is four collectives (two size exchanges, two payloads) - about a millisecond at eight ranks, with every
GPU idle, every round.  All ranks of a ``DistEngine`` on one box can see one POSIX shared-memory segment instead: each rank
owns a slot ``[seq u64 | len u64 | pid u64 | abort u64 | payload]`` (one extra slot carries the aggregator's answer), a
message is published by writing the payload, then the length, then bumping the sequence number, and consumed by polling
the sequence number.  x86 total store order makes payload-before-sequence visible in that order; the poll loop is a
bounded busy wait.

Failure detection: a waiting rank does not rely on the timeout alone.  Every ~thousand polls it looks at the
``abort`` word of every slot (set by ``ShmMailbox.abort`` when a node raised - ``DistEngine.step`` does that before it
re-raises) and checks that the process it is waiting for still exists (``pid`` word, same PID namespace on one box); either
condition raises ``PeerFailure`` within milliseconds instead of leaving the survivors spinning.
"""
import os as _os
import pickle as _pickle
import struct as _struct
import time as _time
from multiprocessing import shared_memory as _shm

_HDR = 32


_OVERSIZE = '__coinn_via_torch_distributed__'


class PeerFailure(RuntimeError):
    """Another rank of the engine raised or died while this one was waiting for its message."""


def _alive(pid):
    try:
        _os.kill(pid, 0)
    except ProcessLookupError:
        return False
    except PermissionError:
        return True
    try:                                               # a zombie (dead, not yet reaped by the launcher) still has a pid
        with open(f'/proc/{pid}/stat') as fp:
            return fp.read().rsplit(')', 1)[1].split()[0] != 'Z'
    except Exception:
        return True


class ShmMailbox:
    """``oversize(obj, peer)`` / ``fetch(peer)``: callbacks that move ONE message over ``torch.distributed`` point-to-point when
    it does not fit a slot (the mailbox then only carries a marker) - set by ``DistEngine``; without them an oversized message
    raises."""

    def __init__(self, name, rank, world, slot_bytes=1 << 20, create=False, timeout_s=1800.0):
        self.rank, self.world, self.slot, self.timeout = rank, world, int(slot_bytes), float(timeout_s)
        self.send_big = self.recv_big = self.bcast_big = None
        size = (world + 1) * self.slot
        self.mem = _shm.SharedMemory(name=name, create=create, size=size if create else 0)
        self.owner = create
        if create:
            for slot in range(world + 1):              # only the headers: pages of /dev/shm are touched when first written
                self.mem.buf[slot * self.slot:slot * self.slot + _HDR] = bytes(_HDR)
        else:
            try:    # attaching processes must not let their resource tracker unlink a segment they do not own
                from multiprocessing import resource_tracker
                resource_tracker.unregister(self.mem._name, 'shared_memory')
            except Exception:
                pass
        _struct.pack_into('<Q', self.mem.buf, self.rank * self.slot + 16, _os.getpid())
        self.check_pids = True
        self.seq = 0           # messages this rank has posted
        self.answers = 0       # aggregator answers consumed / posted

    @property
    def name(self):
        return self.mem.name

    # ------------------------------------------------------------------ raw slots
    def _write(self, slot, seq, obj):
        data = _pickle.dumps(obj, protocol=_pickle.HIGHEST_PROTOCOL)
        big = len(data) + _HDR > self.slot
        if big:
            if self.send_big is None:
                raise ValueError(f'control-plane message of {len(data)} bytes does not fit the {self.slot}-byte mailbox slot')
            data = _pickle.dumps(_OVERSIZE)
        base = slot * self.slot
        self.mem.buf[base + _HDR:base + _HDR + len(data)] = data
        _struct.pack_into('<Q', self.mem.buf, base + 8, len(data))
        _struct.pack_into('<Q', self.mem.buf, base, seq)            # publish last
        return big

    def _read(self, slot, seq):
        base = slot * self.slot
        t0 = spins = 0
        while _struct.unpack_from('<Q', self.mem.buf, base)[0] < seq:
            spins += 1
            if spins & 0x3ff == 0:
                now = _time.monotonic()
                t0 = t0 or now
                if now - t0 > self.timeout:
                    raise TimeoutError(f'control plane: slot {slot} never reached message {seq}')
                self._check_peers(slot, liveness=(spins >> 10) & 0x3f == 1)
                if spins > 200000:
                    _time.sleep(0.0005)                              # long waits (a validation epoch) yield the core
        n = _struct.unpack_from('<Q', self.mem.buf, base + 8)[0]
        return _pickle.loads(bytes(self.mem.buf[base + _HDR:base + _HDR + n]))

    def _check_peers(self, slot, liveness=True):
        """Slow path of a wait on ``slot``: has anybody aborted, is the writer of that slot still there?"""
        for r in range(self.world):
            if _struct.unpack_from('<Q', self.mem.buf, r * self.slot + 24)[0]:
                raise PeerFailure(f'control plane: rank {r} aborted the run')
        if not (liveness and self.check_pids):
            return
        writer = 0 if slot == self.world else slot
        pid = _struct.unpack_from('<Q', self.mem.buf, writer * self.slot + 16)[0]
        if pid and pid != _os.getpid() and not _alive(pid):
            raise PeerFailure(f'control plane: rank {writer} (pid {pid}) is gone')

    def verify_pids(self):
        """Call once every rank has attached: the liveness check is only meaningful when the ranks share a PID namespace.  If
        any recorded pid is not visible from here (ranks in different containers over one /dev/shm) the check is switched
        off for this rank - the abort word and the timeout still apply."""
        for r in range(self.world):
            pid = _struct.unpack_from('<Q', self.mem.buf, r * self.slot + 16)[0]
            if not pid or not _alive(pid):
                self.check_pids = False
        return self.check_pids

    def abort(self):
        """Tell every waiting rank that this one is not going to post again."""
        try:
            _struct.pack_into('<Q', self.mem.buf, self.rank * self.slot + 24, 1)
        except Exception:
            pass

    # ------------------------------------------------------------------ round protocol
    def gather(self, obj):
        """Every rank posts ``obj``; rank 0 returns the list of all ranks' objects, the others ``None``."""
        self.seq += 1
        if self._write(self.rank, self.seq, obj):
            if self.rank != 0:
                self.send_big(obj, 0)                  # the marker is posted: rank 0 will ask for the payload point-to-point
        if self.rank != 0:
            return None
        out = []
        for r in range(self.world):
            got = self._read(r, self.seq)
            if isinstance(got, str) and got == _OVERSIZE:
                got = obj if r == 0 else self.recv_big(r)
            out.append(got)
        return out

    def broadcast(self, obj=None):
        """Rank 0 posts ``obj`` in the answer slot; every rank returns it."""
        self.answers += 1
        if self.rank == 0:
            if self._write(self.world, self.answers, obj):
                self.bcast_big(obj)
            return obj
        got = self._read(self.world, self.answers)
        if isinstance(got, str) and got == _OVERSIZE:
            got = self.bcast_big(None)
        return got

    def close(self):
        try:
            self.mem.close()
            if self.owner:
                self.mem.unlink()
        except Exception:
            pass
