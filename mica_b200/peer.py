"""Peer-memory group for the multi-GPU histogram exchange.

One small buffer per rank, mapped by every other rank of the box through CUDA IPC, so that
``mica_select_peer_reduce`` can publish / signal / wait / sum over NVLink inside one kernel
instead of an NCCL all-reduce per radix round.  ``torch.distributed`` is only used once, to
hand the 64-byte IPC handles around.  ``PeerHistogram.emulate`` builds the same structure for
several "ranks" inside one process (plain pointers, no IPC) for the single-GPU tests."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import lib, check


class PeerHistogram:
    def __init__(self, device, rank: int, world: int, group=None, _emulated=None):
        self.device, self.rank, self.world = torch.device(device), int(rank), int(world)
        self.epoch = 0
        self._opened = []
        if _emulated is not None:                       # in-process emulation: pointers are shared directly
            self._own = None
            ptrs = _emulated
        else:
            import torch.distributed as dist
            torch.cuda.set_device(self.device)
            own = C.c_void_p()
            handle = (C.c_ubyte * 64)()
            check(lib.mica_peer_alloc(C.byref(own), handle), 'peer_alloc')
            self._own = own.value
            gathered = [None] * self.world
            dist.all_gather_object(gathered, bytes(handle), group=group)
            ptrs = []
            for r, h in enumerate(gathered):
                if r == self.rank:
                    ptrs.append(self._own)
                    continue
                p = C.c_void_p()
                buf = (C.c_ubyte * 64).from_buffer_copy(h)
                check(lib.mica_peer_open(buf, C.byref(p)), f'peer_open(rank {r})')
                self._opened.append(p.value)
                ptrs.append(p.value)
        self.table = torch.tensor(ptrs, dtype=torch.int64, device=self.device)

    @classmethod
    def emulate(cls, device, world: int):
        """``world`` groups for ranks 0..world-1 living in this process (tests)."""
        device = torch.device(device)
        torch.cuda.set_device(device)
        ptrs = []
        for _ in range(world):
            p = C.c_void_p()
            check(lib.mica_peer_alloc(C.byref(p), None), 'peer_alloc')
            ptrs.append(p.value)
        groups = [cls(device, r, world, _emulated=ptrs) for r in range(world)]
        groups[0]._owned_all = ptrs                     # freed with the first group
        return groups

    def reduce(self, stats, round_index: int, stream=None):
        """Sum every rank's histogram of this round into ``stats`` (an ops.OrderStats), in stream order."""
        # the slot alternates with the EPOCH (not the select step): a map has an odd number of steps, so the
        # last exchange of one map and the first of the next would otherwise share a slot back to back
        self.epoch += 1
        st = C.c_void_p(stream if stream is not None else torch.cuda.current_stream(self.device).cuda_stream)
        check(lib.mica_select_peer_reduce(stats._p, C.c_void_p(self.table.data_ptr()), self.rank, self.world,
                                          self.epoch & 1, self.epoch, st), 'select_peer_reduce')

    def close(self):
        torch.cuda.synchronize(self.device)
        for p in self._opened:
            lib.mica_peer_close(C.c_void_p(p))
        self._opened = []
        for p in getattr(self, '_owned_all', []) + ([self._own] if self._own else []):
            lib.mica_peer_free(C.c_void_p(p))
        self._own, self._owned_all = None, []
