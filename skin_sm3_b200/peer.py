"""Peer-memory (NVLink / NVSwitch) exchange for the row-sharded objective: every rank stores its rows straight into
every rank's copy of the global buffers -- no NCCL on the hot path.

Buffers live in torch symmetric memory (``torch.distributed._symmetric_memory``), which hands back the peer-mapped
device pointers and a stream-ordered cross-rank barrier (signal pads).  Two buffers per slot:
  z_cols [2*n_global, D] bf16   normalised embeddings in the reference's [all first ; all second] order
  stats  [2*n_global, 4] fp32   (g_pos, g_lse, neg_sum, -) per global row for the row-local symmetric backward
Slots are double-buffered: a rank that has passed both barriers of step k+1 knows every peer finished reading the
slot of step k, so slot (k+2) % 2 can be overwritten.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Tuple

import torch
import torch.distributed as dist

from ._lib import check, lib, ptr, stream_ptr

_CACHE: Dict[Tuple, "PeerBuffers"] = {}


class PeerUnavailable(RuntimeError):
    pass


class PeerBuffers:
    DEPTH = 2

    def __init__(self, group, n_global: int, d: int, device: torch.device, depth: int = 0):
        if depth:
            self.DEPTH = int(depth)            # slots kept alive between a forward and its later backward (cal_logits)
        try:
            import torch.distributed._symmetric_memory as symm
        except Exception as e:  # pragma: no cover
            raise PeerUnavailable(f"torch symmetric memory is not importable: {e!r}") from e
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        if self.world > 16:
            raise PeerUnavailable("peer exchange supports up to 16 ranks")
        self.n_global, self.d = n_global, d
        self.z, self.zh, self.zp = [], [], []
        self.st, self.sth, self.stp = [], [], []
        try:
            for _ in range(self.DEPTH):
                z = symm.empty((2 * n_global, d), dtype=torch.bfloat16, device=device)
                zh = symm.rendezvous(z, self.group)
                st = symm.empty((2 * n_global, 4), dtype=torch.float32, device=device)
                sth = symm.rendezvous(st, self.group)
                self.z.append(z); self.zh.append(zh); self.zp.append((C.c_void_p * self.world)(*zh.buffer_ptrs))
                self.st.append(st); self.sth.append(sth); self.stp.append((C.c_void_p * self.world)(*sth.buffer_ptrs))
        except PeerUnavailable:
            raise
        except Exception as e:
            raise PeerUnavailable(f"symmetric-memory rendezvous failed: {e!r}") from e
        # NVSwitch multicast mapping of the same buffers (0 when the fabric / driver has no multicast support)
        self.zmc = [int(getattr(h, "multicast_ptr", 0) or 0) for h in self.zh]
        self.stmc = [int(getattr(h, "multicast_ptr", 0) or 0) for h in self.sth]
        # Measured on 8 x B200 (cfg4, 4 MiB of rows per rank): unicast stores 54 us vs multicast 44 us for the scatter
        # kernel, but the step is faster with unicast (0.759 vs 0.819 ms) -- the inbound 28 MiB per GPU is the floor
        # either way -- so multicast is opt-in (SM3_PEER_MULTICAST=1).
        want_mc = os.environ.get("SM3_PEER_MULTICAST", "0") == "1"
        self.multicast = want_mc and all(self.zmc) and all(self.stmc)
        # split barrier (signal / wait) on a zero-initialised symmetric flag buffer + a side stream for the exchange
        try:
            # 64 flag slots (channel * 16 + source rank) + local ticket words of the fused exchange kernels
            self.flags = symm.empty((128,), dtype=torch.int32, device=device)
            self.flags.zero_()
            self.fh = symm.rendezvous(self.flags, self.group)
            self.fp = (C.c_void_p * self.world)(*self.fh.buffer_ptrs)
            self.fh.barrier(channel=2)           # nobody signals before everybody has zeroed its flags
        except Exception as e:
            raise PeerUnavailable(f"symmetric-memory flag buffer failed: {e!r}") from e
        self.side = torch.cuda.Stream(device=device)
        self.step = 0

    # ---- split barrier ----
    def signal(self, channel: int, epoch: int) -> None:
        check(lib().sm3_peer_signal(self.fp, self.world, self.rank, channel, epoch & 0x7FFFFFFF, stream_ptr()),
              "sm3_peer_signal")

    def wait(self, channel: int, epoch: int) -> None:
        check(lib().sm3_peer_wait(ptr(self.flags), self.world, channel, epoch & 0x7FFFFFFF, stream_ptr()),
              "sm3_peer_wait")

    def store_z(self, slot: int, z_local: torch.Tensor, n_local: int) -> None:
        """scatter kernel only (no barrier): the caller brackets it with signal()/wait()."""
        with torch.cuda.device(z_local.device):
            if self.multicast:
                check(lib().sm3_peer_multicast_rows(ptr(z_local), n_local, self.rank * n_local, self.n_global,
                                                    self.d * 2, self.zmc[slot], stream_ptr()), "sm3_peer_multicast_rows")
            else:
                check(lib().sm3_peer_scatter_rows(ptr(z_local), n_local, self.rank * n_local, self.n_global, self.d * 2,
                                                  self.zp[slot], self.world, stream_ptr()), "sm3_peer_scatter_rows")

    def store_stats(self, slot: int, g_pos, g_lse, nsum, n_local: int) -> None:
        with torch.cuda.device(g_pos.device):
            if self.multicast:
                check(lib().sm3_peer_multicast_stats(ptr(g_pos), ptr(g_lse), ptr(nsum), n_local, self.rank * n_local,
                                                     self.n_global, self.stmc[slot], stream_ptr()),
                      "sm3_peer_multicast_stats")
            else:
                check(lib().sm3_peer_scatter_stats(ptr(g_pos), ptr(g_lse), ptr(nsum), n_local, self.rank * n_local,
                                                   self.n_global, self.stp[slot], self.world, stream_ptr()),
                      "sm3_peer_scatter_stats")

    def next_slot(self) -> int:
        k = self.step % self.DEPTH
        self.step += 1
        return k

    # ---- forward exchange: rows of the local normalised z -> every rank's z_cols[slot] ----
    def scatter_z(self, slot: int, z_local: torch.Tensor, n_local: int) -> torch.Tensor:
        assert z_local.dtype == torch.bfloat16 and z_local.is_contiguous() and z_local.shape == (2 * n_local, self.d)
        from .functional import _mark
        with torch.cuda.device(z_local.device):
            if self.multicast:
                check(lib().sm3_peer_multicast_rows(ptr(z_local), n_local, self.rank * n_local, self.n_global,
                                                    self.d * 2, self.zmc[slot], stream_ptr()), "sm3_peer_multicast_rows")
            else:
                check(lib().sm3_peer_scatter_rows(ptr(z_local), n_local, self.rank * n_local, self.n_global, self.d * 2,
                                                  self.zp[slot], self.world, stream_ptr()), "sm3_peer_scatter_rows")
            _mark("scatter_z_kernel")
            self.zh[slot].barrier(channel=0)
        return self.z[slot]

    # ---- backward exchange: (g_pos, g_lse, neg_sum) of the local rows -> every rank's stats[slot] ----
    def scatter_stats(self, slot: int, g_pos, g_lse, nsum, n_local: int) -> torch.Tensor:
        with torch.cuda.device(g_pos.device):
            if self.multicast:
                check(lib().sm3_peer_multicast_stats(ptr(g_pos), ptr(g_lse), ptr(nsum), n_local, self.rank * n_local,
                                                     self.n_global, self.stmc[slot], stream_ptr()),
                      "sm3_peer_multicast_stats")
            else:
                check(lib().sm3_peer_scatter_stats(ptr(g_pos), ptr(g_lse), ptr(nsum), n_local, self.rank * n_local,
                                                   self.n_global, self.stp[slot], self.world, stream_ptr()),
                      "sm3_peer_scatter_stats")
            self.sth[slot].barrier(channel=1)
        return self.st[slot]


def get_peer_buffers(group, n_global: int, d: int, device: torch.device, depth: int = 0) -> PeerBuffers:
    """depth = 0: the two-slot buffers of the fused step; depth > 0: a separate pool whose slots stay untouched between a
    term's forward and its backward (the logits form, where autograd runs the backward later)."""
    g = group if group is not None else dist.group.WORLD
    key = (id(g), n_global, d, device.index, depth)
    pb = _CACHE.get(key)
    if pb is None:
        pb = PeerBuffers(g, n_global, d, device, depth)
        _CACHE[key] = pb
    return pb
