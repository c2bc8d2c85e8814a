"""Device-resident descriptor database shards (keyframe database / local map) and the multi-GPU
top-2 merge.

The reference keeps every keyframe's descriptor Mats in host memory (include/keyFrame.h:69-96) and
matches one candidate at a time (src/mapHandler.cpp:3301-3409).  Here the database rows live in HBM,
row-sharded over the ranks of a torch.distributed job (one process per GPU); a query is matched
against the local shard by the brute-force kernel, the per-query packed top-2 keys of all shards are
all-gathered over NCCL/NVLink and merged lexicographically, so the G-GPU result is bit-identical to
the single-GPU one (lowest global index wins ties).  torch is plumbing only: device memory, streams
and the collective.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib as L
from .matching import Context


def _ptr(t: torch.Tensor) -> C.c_void_p:
    return C.c_void_p(t.data_ptr())


def shard_bounds(n_rows: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous row ranges [g*ceil(N/G), ...)."""
    per = (n_rows + world - 1) // world
    lo = min(n_rows, rank * per)
    return lo, min(n_rows, lo + per)


class DeviceOps:
    """Thin wrappers of the plm_dev_* entry points on torch CUDA tensors (current torch stream)."""

    def __init__(self, device: Optional[int] = None):
        self.device = torch.cuda.current_device() if device is None else device
        self.ctx = Context(self.device)
        self.lib = L.load()

    def _bind_stream(self):
        self.ctx.set_stream(torch.cuda.current_stream(self.device).cuda_stream)

    def knn2(self, q: torch.Tensor, db: torch.Tensor, idx_base: int = 0, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """q: nq x 32 uint8, db: n x 32 uint8 (both CUDA, contiguous) -> nq x 2 int64 packed keys."""
        assert q.is_cuda and db.is_cuda and q.dtype == torch.uint8 and db.dtype == torch.uint8
        assert q.is_contiguous() and db.is_contiguous() and q.shape[-1] == 32 and db.shape[-1] == 32
        nq, n = q.shape[0], db.shape[0]
        if out is None:
            out = torch.empty((nq, 2), dtype=torch.int64, device=q.device)
        self._bind_stream()
        L.check(self.lib.plm_dev_knn2(self.ctx.handle, _ptr(q), nq, _ptr(db), n, idx_base, _ptr(out)), "plm_dev_knn2")
        return out

    def top2_merge(self, parts: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """parts: P x nq x 2 int64 packed keys -> nq x 2 (two smallest keys per query)."""
        assert parts.is_cuda and parts.dtype == torch.int64 and parts.is_contiguous() and parts.dim() == 3
        p, nq = parts.shape[0], parts.shape[1]
        if out is None:
            out = torch.empty((nq, 2), dtype=torch.int64, device=parts.device)
        self._bind_stream()
        L.check(self.lib.plm_dev_top2_merge(self.ctx.handle, _ptr(parts), p, nq, _ptr(out)), "plm_dev_top2_merge")
        return out

    def nnr_accept(self, top2: torch.Tensor, nnr: float, m12: torch.Tensor, count: torch.Tensor) -> None:
        assert m12.dtype == torch.int32 and count.dtype == torch.int32
        self._bind_stream()
        L.check(self.lib.plm_dev_nnr_accept(self.ctx.handle, _ptr(top2), top2.shape[0], C.c_float(nnr), _ptr(m12),
                                            _ptr(count)), "plm_dev_nnr_accept")

    def cross_check(self, m12: torch.Tensor, i1_base: int, m21: torch.Tensor, count: torch.Tensor) -> None:
        self._bind_stream()
        L.check(self.lib.plm_dev_cross_check(self.ctx.handle, _ptr(m12), m12.shape[0], i1_base, _ptr(m21),
                                             m21.shape[0], _ptr(count)), "plm_dev_cross_check")


class ShardedDescriptorDB:
    """Row-sharded descriptor database over the ranks of the default process group.

    ``rows`` (host, n x 32 uint8) is the WHOLE database on every rank for construction convenience;
    each rank uploads only its contiguous shard.  With world == 1 no collective is issued.
    """

    def __init__(self, rows: Optional[np.ndarray] = None, n_rows: Optional[int] = None, device: Optional[int] = None,
                 group=None, shard: Optional[torch.Tensor] = None):
        import torch.distributed as dist
        self.dist = dist if (dist.is_available() and dist.is_initialized()) else None
        self.group = group
        self.world = self.dist.get_world_size(group) if self.dist else 1
        self.rank = self.dist.get_rank(group) if self.dist else 0
        self.ops = DeviceOps(device)
        dev = torch.device("cuda", self.ops.device)
        if shard is not None:
            assert n_rows is not None
            self.n_rows = n_rows
            self.lo, self.hi = shard_bounds(n_rows, self.world, self.rank)
            assert shard.shape[0] == self.hi - self.lo
            self.shard = shard
        else:
            self.n_rows = int(rows.shape[0])
            self.lo, self.hi = shard_bounds(self.n_rows, self.world, self.rank)
            self.shard = torch.from_numpy(np.ascontiguousarray(rows[self.lo:self.hi])).to(dev)
        self._gather = None

    def knn2_local(self, q: torch.Tensor) -> torch.Tensor:
        """Packed top-2 of the queries over THIS rank's shard, with global row indices."""
        return self.ops.knn2(q, self.shard, idx_base=self.lo)

    def knn2(self, q: torch.Tensor) -> torch.Tensor:
        """Packed top-2 over the whole database: local kernel -> all_gather (NCCL) -> merge kernel."""
        local = self.knn2_local(q)
        if self.world == 1:
            return local
        nq = q.shape[0]
        if self._gather is None or self._gather.shape[1] != nq:
            self._gather = torch.empty((self.world, nq, 2), dtype=torch.int64, device=q.device)
        self.dist.all_gather_into_tensor(self._gather, local, group=self.group)
        return self.ops.top2_merge(self._gather)

    def match_nnr(self, q: torch.Tensor, nnr: float):
        """StVO::matchNNR of the queries against the whole (sharded) database -> (count, m12)."""
        top2 = self.knn2(q)
        m12 = torch.full((q.shape[0],), -1, dtype=torch.int32, device=q.device)
        count = torch.zeros(1, dtype=torch.int32, device=q.device)
        self.ops.nnr_accept(top2, nnr, m12, count)
        return count, m12
