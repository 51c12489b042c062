"""Symmetric (peer-mapped) device buffers: the memory the fused NVLink kernels read and write.

A ``SymmetricBuffer`` is the same-sized allocation on every rank, mapped into every peer's address
space (CUDA VMM through ``torch.distributed._symmetric_memory``), optionally with an NVLS multicast
alias.  The process group is only needed here, for the rendezvous - the per-step data plane never
calls into NCCL.  With ``world == 1`` or on CPU it is just a local tensor.
"""
import torch as _torch
import torch.distributed as _dist


def _world(group=None):
    return _dist.get_world_size(group) if _dist.is_available() and _dist.is_initialized() else 1


def _rank(group=None):
    return _dist.get_rank(group) if _dist.is_available() and _dist.is_initialized() else 0


class SymmetricBuffer:
    def __init__(self, numel, dtype, device, group=None, zero=True):
        self.numel, self.dtype, self.device, self.group = int(numel), dtype, _torch.device(device), group
        self.world, self.rank = _world(group), _rank(group)
        self.handle = None
        self.multicast_ptr = 0
        if self.world > 1 and self.device.type == 'cuda':
            import torch.distributed._symmetric_memory as symm
            grp = group if group is not None else _dist.group.WORLD
            try:
                symm.enable_symm_mem_for_group(grp.group_name)
            except Exception:  # deprecated / not needed on new versions
                pass
            self.local = symm.empty(self.numel, dtype=dtype, device=self.device)
            self.handle = symm.rendezvous(self.local, grp)
            self.peer_ptrs = [int(p) for p in self.handle.buffer_ptrs]
            try:
                self.multicast_ptr = int(self.handle.multicast_ptr or 0)
            except Exception:
                self.multicast_ptr = 0
        else:
            self.local = _torch.empty(self.numel, dtype=dtype, device=self.device)
            self.peer_ptrs = [self.local.data_ptr()] if self.device.type == 'cuda' else [0]
        if zero:
            self.local.zero_()
        if self.handle is not None:
            _torch.cuda.synchronize(self.device)
            self.handle.barrier()

    @property
    def is_symmetric(self):
        return self.handle is not None

    def peer_tensor(self, peer):
        """A tensor aliasing ``peer``'s buffer (debug / tests)."""
        if self.handle is None:
            return self.local
        return self.handle.get_buffer(peer, (self.numel,), self.dtype)

    def barrier(self):
        if self.handle is not None:
            self.handle.barrier()
