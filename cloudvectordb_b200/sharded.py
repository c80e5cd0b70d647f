"""Shard/merge layer: the database is row-sharded over the ranks of a
``torch.distributed`` group (one process per GPU); every rank searches its
shard with the fused kernel, the per-rank candidates are exchanged with ONE
all-gather (NCCL over NVLink; [nq, k] 64-bit keys per rank) and each rank does
the final k-way select (a head-pointer merge of the sorted per-rank lists).

Top-k over a union of row sets == top-k of the per-set top-k lists, so this is
the only collective on the path.  Global ids are the
concatenation of the shards in rank order.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch
import torch.distributed as dist

from .index import IndexFlat, merge_topk


def shard_bounds(n: int, world: int, rank: int):
    """Contiguous row range [lo, hi) of `rank` when n rows are split over `world` ranks."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class ShardedIndex:
    def __init__(self, d: int, metric: str = "ip", storage: str = "bf16", device: Optional[int] = None,
                 group=None, local_index=None):
        """`local_index` exists so the host logic can be exercised on CPU (gloo) with a stand-in that offers
        the same search / search_keys / merge_keys surface; the product path builds an IndexFlat on `device`."""
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.d, self.metric, self.storage = int(d), metric.lower(), storage.lower()
        self.device_path = local_index is None
        if local_index is None:
            if device is None:
                device = torch.cuda.current_device()
            local_index = IndexFlat(d, metric, storage, device)
        self.device = device
        self.local = local_index
        self.id_base = 0
        self._ntotal = 0
        self._counts = [0] * self.world

    # ------------------------------------------------------------------ add
    def add(self, x) -> None:
        """x is the SAME full matrix on every rank; each rank keeps its contiguous slice.
        Only valid on an empty index (global ids stay the row numbers of x)."""
        if self._ntotal != 0:
            raise ValueError("add() of a replicated matrix needs an empty index; use add_local() to append shards")
        lo, hi = shard_bounds(int(x.shape[0]), self.world, self.rank)
        self.add_local(x[lo:hi])

    def add_local(self, x_shard) -> None:
        """Every rank appends its own rows; ids are assigned in rank order."""
        self.local.add(x_shard)
        n_loc = int(self.local.ntotal)
        if self.world > 1:
            counts = [None] * self.world
            dist.all_gather_object(counts, n_loc, group=self.group)
        else:
            counts = [n_loc]
        self._counts = [int(c) for c in counts]
        self.id_base = sum(self._counts[: self.rank])
        self._ntotal = sum(self._counts)

    @property
    def ntotal(self) -> int:
        return self._ntotal

    def reset(self) -> None:
        self.local.reset()
        self.id_base, self._ntotal, self._counts = 0, 0, [0] * self.world

    def set_groups_local(self, group_shard) -> None:
        self.local.set_groups(group_shard)

    # --------------------------------------------------------------- search
    def search(self, q, k: int, *, self_ids=None, group_q=None, profile: bool = False):
        """q is replicated on every rank.  self_ids are GLOBAL row ids.  Host queries
        are staged to the rank's GPU (the exchange runs over NCCL) and the merged
        result is returned on the host."""
        host_io = False
        if self.device_path and self.world > 1 and not (isinstance(q, torch.Tensor) and q.is_cuda):
            host_io = True
            q = torch.as_tensor(q).to(f"cuda:{self.device}", non_blocking=True)
        kw = {"profile": profile} if self.device_path else {}
        local_self = None
        if self_ids is not None:
            s = torch.as_tensor(self_ids).to(torch.int64)
            in_shard = (s >= self.id_base) & (s < self.id_base + self._counts[self.rank])
            local_self = torch.where(in_shard, s - self.id_base, torch.full_like(s, -1)).to(torch.int32)
        if self.world == 1:
            return self.local.search(q, k, self_ids=local_self, group_q=group_q, id_base=self.id_base, **kw)
        # one exchange step, in the kernel's own 64-bit candidate keys: every rank contributes its sorted top-k
        # (8 bytes per candidate, ids already global) to ONE all_gather_into_tensor, then merges the `world`
        # sorted lists per query with a head-pointer merge
        keys = self.local.search_keys(q, k, self_ids=local_self, group_q=group_q, id_base=self.id_base, **kw)
        gathered = torch.empty((self.world,) + tuple(keys.shape), dtype=keys.dtype, device=keys.device)
        dist.all_gather_into_tensor(gathered.view(-1, keys.shape[-1]), keys.contiguous(), group=self.group)
        Dm, Im = self.local.merge_keys(gathered, k)
        if host_io:
            return Dm.cpu(), Im.cpu()
        return Dm, Im


class ShardedIVFFlat:
    """IVF-Flat over row shards: every rank holds the same coarse quantizer (centroids) and the inverted
    lists of ITS rows; a search probes the same lists on every rank, scans the local parts, and merges the
    per-rank candidates like ShardedIndex (one all-gather + k-way select).  Global ids are the
    concatenation of the shards in rank order."""

    def __init__(self, d: int, nlist: int, metric: str = "l2", device: Optional[int] = None, group=None):
        from .ivf import IndexIVFFlat
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        if device is None:
            device = torch.cuda.current_device()
        self.device, self.metric, self.nlist = device, metric.lower(), int(nlist)
        self.local = IndexIVFFlat(d, nlist, metric, device)
        self.nprobe = 1
        self.id_base, self._ntotal = 0, 0

    def train(self, x_local, niter: int = 10, seed: int = 42, centroids=None) -> None:
        """Distributed k-means over the ranks' points (sums / counts all-reduced), or adopt given centroids."""
        if centroids is None:
            from .kmeans import Kmeans
            km = Kmeans(self.local.d, self.nlist, niter=niter, seed=seed, device=self.device, group=self.group)
            km.train(self.local._to_device(x_local))
            centroids = km.centroids
        self.local.train(None, centroids=centroids)

    def add_local(self, x_shard) -> None:
        self.local.add(x_shard)
        n_loc = int(self.local.ntotal)
        counts = [n_loc]
        if self.world > 1:
            counts = [None] * self.world
            dist.all_gather_object(counts, n_loc, group=self.group)
        self.id_base = sum(int(c) for c in counts[: self.rank])
        self._ntotal = sum(int(c) for c in counts)

    @property
    def ntotal(self) -> int:
        return self._ntotal

    def search(self, q, k: int, nprobe: Optional[int] = None):
        qd = self.local._to_device(q)
        D, I = self.local.search(qd, k, nprobe=nprobe or self.nprobe)
        I = torch.where(I >= 0, I + self.id_base, I)
        if self.world == 1:
            return D, I
        Dg = torch.empty((self.world,) + tuple(D.shape), dtype=D.dtype, device=D.device)
        Ig = torch.empty((self.world,) + tuple(I.shape), dtype=I.dtype, device=I.device)
        dist.all_gather_into_tensor(Dg.view(-1, D.shape[-1]), D.contiguous(), group=self.group)
        dist.all_gather_into_tensor(Ig.view(-1, I.shape[-1]), I.contiguous(), group=self.group)
        return merge_topk(Dg, Ig, k, self.metric)

    def close(self) -> None:
        self.local.close()
