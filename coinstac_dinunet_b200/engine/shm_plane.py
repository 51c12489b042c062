"""Shared-memory mailbox for the JSON control plane of a single-node ``DistEngine``.

Per engine round every site sends one small dict to the aggregator and gets one back.  Over ``torch.distributed`` object
collectives that is four collectives (two size exchanges, two payloads) - about a millisecond at eight ranks, with every
GPU idle, every round.  All ranks of a ``DistEngine`` on one box can see one POSIX shared-memory segment instead: each rank
owns a slot ``[seq u64 | len u64 | payload]`` (one extra slot carries the aggregator's answer), a message is published by
writing the payload, then the length, then bumping the sequence number, and consumed by polling the sequence number.  x86
total store order makes payload-before-sequence visible in that order; the poll loop is a bounded busy wait.
"""
import pickle as _pickle
import struct as _struct
import time as _time
from multiprocessing import shared_memory as _shm

_HDR = 16


_OVERSIZE = '__coinn_via_torch_distributed__'


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
                if spins > 200000:
                    _time.sleep(0.0005)                              # long waits (a validation epoch) yield the core
        n = _struct.unpack_from('<Q', self.mem.buf, base + 8)[0]
        return _pickle.loads(bytes(self.mem.buf[base + _HDR:base + _HDR + n]))

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
