"""IVF coarse-quantizer k-means (README.md:2 "building the vectordb"): Lloyd
iterations whose assignment step is the fused GEMM + top-1 kernel (centroids
are the index rows, the points are the queries) and whose update step is a
scatter-add kernel, all-reduced across ranks when torch.distributed is up.

FAISS ``Kmeans(d, k, niter, seed)`` shape: ``train(x)``, ``centroids``,
``assign(x)``, ``obj``.  Empty clusters take half of the largest cluster (FAISS
splits a big cluster too; here the donor is the largest one, so that every rank
of a sharded run re-seeds identically); ``split_empty=False`` keeps the previous
centroid instead.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch
import torch.distributed as dist

from . import _C
from .index import IndexFlat


class Kmeans:
    def __init__(self, d: int, k: int, niter: int = 10, seed: int = 42, storage: str = "bf16",
                 device: Optional[int] = None, group=None, split_empty: bool = True, split_eps: float = 1.0 / 1024):
        self.d, self.k, self.niter, self.seed = int(d), int(k), int(niter), int(seed)
        self.storage = storage
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.group = group
        self.split_empty, self.split_eps = bool(split_empty), float(split_eps)
        self.last_nsplit: Optional[torch.Tensor] = None   # device int32: clusters re-seeded by the last step
        self.centroids: Optional[torch.Tensor] = None
        self.obj = []
        self._index: Optional[IndexFlat] = None

    def _dev(self):
        return torch.device(f"cuda:{self.device}")

    def _to_device(self, x):
        if isinstance(x, np.ndarray):
            x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
        if x.dtype not in (torch.float32, torch.bfloat16, torch.float16):
            x = x.float()
        return x.to(self._dev()).contiguous()

    def init_centroids(self, x: torch.Tensor) -> torch.Tensor:
        """k distinct local points chosen with `seed` (rank 0's choice is broadcast).  Every rank checks the
        SAME condition (rank 0's point count, shared by a collective), so either all raise or none does."""
        n0 = torch.tensor([int(x.shape[0])], dtype=torch.int64, device=x.device)
        if dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.broadcast(n0, src=dist.get_global_rank(self.group, 0) if self.group is not None else 0,
                           group=self.group)
        if int(n0.item()) < self.k:
            raise ValueError(f"k-means with k={self.k} needs at least k points on the initialising rank "
                             f"(it holds {int(n0.item())}); pass init_centroids= or fewer clusters")
        g = torch.Generator(device="cpu").manual_seed(self.seed)
        perm = torch.randperm(x.shape[0], generator=g)[: self.k].to(x.device)
        c = x[perm].float().contiguous()
        if dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.broadcast(c, src=dist.get_global_rank(self.group, 0) if self.group is not None else 0,
                           group=self.group)
        return c

    def _set_centroids(self, c: torch.Tensor) -> None:
        if self._index is None:
            self._index = IndexFlat(self.d, "l2", self.storage, self.device)
        self._index.reset()
        self._index.add(c)

    def step(self, x: torch.Tensor, profile: bool = False):
        """One Lloyd iteration on this rank's points: returns (assign, objective).
        With profile=True, self.last_timing holds the device time (ms) of each phase."""
        lib = _C.lib()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if profile else None
        if profile:
            ev[0].record()
        self._set_centroids(self.centroids)
        assign, dist_sq = self._index.assign(x, return_dist=True)
        if profile:
            ev[1].record()
        sums = torch.zeros((self.k, self.d), dtype=torch.float32, device=x.device)
        counts = torch.zeros((self.k,), dtype=torch.int32, device=x.device)
        stream = int(torch.cuda.current_stream(self.device).cuda_stream)
        dt = {torch.float32: _C.DTYPE_F32, torch.bfloat16: _C.DTYPE_BF16, torch.float16: _C.DTYPE_F16}[x.dtype]
        _C.check(lib.cvdb_kmeans_accumulate(x.data_ptr(), x.shape[0], self.d, dt, assign.data_ptr(), sums.data_ptr(),
                                            counts.data_ptr(), stream))
        obj = dist_sq.double().sum()
        if profile:
            ev[2].record()
        if dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_reduce(sums, group=self.group)
            dist.all_reduce(counts, group=self.group)
            dist.all_reduce(obj, group=self.group)
        _C.check(lib.cvdb_kmeans_finalize(sums.data_ptr(), counts.data_ptr(), self.k, self.d,
                                          self.centroids.data_ptr(), stream))
        if self.split_empty:
            self.last_nsplit = torch.zeros((1,), dtype=torch.int32, device=x.device)
            _C.check(lib.cvdb_kmeans_split_empty(self.centroids.data_ptr(), counts.data_ptr(), self.k, self.d,
                                                 self.split_eps, self.last_nsplit.data_ptr(), stream))
        if profile:
            ev[3].record()
            torch.cuda.synchronize(self.device)
            self.last_timing = {"assign_ms": ev[0].elapsed_time(ev[1]), "update_ms": ev[1].elapsed_time(ev[2]),
                                "reduce_finalize_ms": ev[2].elapsed_time(ev[3])}
        self.last_counts = counts
        return assign, obj

    def train(self, x, init_centroids=None):
        x = self._to_device(x)
        if x.shape[1] != self.d:
            raise ValueError("x has the wrong dimension")
        c = self.init_centroids(x) if init_centroids is None else self._to_device(init_centroids).float()
        self.centroids = c.clone()
        self.obj = []
        for _ in range(self.niter):
            _, obj = self.step(x)
            self.obj.append(float(obj))
        self._set_centroids(self.centroids)
        return self.obj[-1] if self.obj else None

    def assign(self, x):
        x = self._to_device(x)
        if self._index is None or self._index.ntotal != self.k:
            self._set_centroids(self.centroids)
        return self._index.assign(x, return_dist=True)
