"""IVF-Flat on top of the coarse quantizer (README.md:2
"building the vectordb"): k-means centroids, inverted lists built from the k=1
assignment, and an nprobe search that scans only the probed lists.

FAISS ``IndexIVFFlat`` shape: ``train(x)``, ``add(x)``, ``search(q, k)`` with the
``nprobe`` attribute, ``ntotal``, ``nlist``.  Everything numeric runs in the CUDA
library: the coarse search and the assignment are the fused GEMM + top-k kernel
over the centroids, the rows are bucketed into lists by histogram / scan /
scatter kernels, and the list scan is the grouped variant of the fused kernel
(one work item per probed list and up to 128 of the queries probing it).
torch only provides the device buffers.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import _C
from .index import IndexFlat, _Buf
from .kmeans import Kmeans


class IndexIVFFlat:
    def __init__(self, d: int, nlist: int, metric: str = "l2", device: int = 0):
        self.d, self.nlist, self.metric, self.device = int(d), int(nlist), metric.lower(), int(device)
        self.nprobe = 1
        self.quantizer = IndexFlat(d, self.metric, "bf16", device)   # the centroids
        self.lists = IndexFlat(d, self.metric, "bf16", device)       # the rows, stored list-major once grouped
        self.centroids: Optional[torch.Tensor] = None
        self._assign = torch.empty((0,), dtype=torch.int32, device=self._dev())  # list of every row, by row id
        self._grouped = False
        self.list_offsets: Optional[torch.Tensor] = None

    def _dev(self):
        return torch.device(f"cuda:{self.device}")

    def _to_device(self, x):
        if isinstance(x, np.ndarray):
            x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
        if x.dtype not in (torch.float32, torch.bfloat16, torch.float16):
            x = x.float()
        return x.to(self._dev()).contiguous()

    # ------------------------------------------------------------------ build
    @property
    def is_trained(self) -> bool:
        return self.centroids is not None

    @property
    def ntotal(self) -> int:
        return self.lists.ntotal

    def train(self, x, niter: int = 10, seed: int = 42, centroids=None) -> None:
        """k-means coarse quantizer (or adopt given centroids)."""
        if centroids is None:
            km = Kmeans(self.d, self.nlist, niter=niter, seed=seed, device=self.device)
            km.train(self._to_device(x))
            centroids = km.centroids
        self.centroids = self._to_device(centroids).float().contiguous()
        if self.centroids.shape != (self.nlist, self.d):
            raise ValueError("centroids must be [nlist, d]")
        self.quantizer.reset()
        self.quantizer.add(self.centroids)

    def assign_lists(self, x) -> torch.Tensor:
        """List of every row: the best centroid under the index metric (int32, on the GPU)."""
        _, I = self.quantizer.search(self._to_device(x), 1)
        return I[:, 0].to(torch.int32)

    def add(self, x) -> None:
        if not self.is_trained:
            raise RuntimeError("train() first")
        x = self._to_device(x)
        a = self.assign_lists(x)
        self.lists.add(x)                # add() un-groups the stored rows (new rows are appended in insertion order)
        self._assign = torch.cat([self._assign, a])
        self._grouped = False

    def _group(self) -> None:
        """Store the rows list-major (histogram / scan / scatter kernels in the library)."""
        stream = int(torch.cuda.current_stream(self.device).cuda_stream)
        self._assign = self._assign.contiguous()
        _C.check(_C.lib().cvdb_index_group_by_list(self.lists._h, self._assign.data_ptr(), self.nlist, stream))
        self.list_offsets = torch.empty((self.nlist + 1,), dtype=torch.int32, device=self._dev())
        _C.check(_C.lib().cvdb_index_list_offsets(self.lists._h, self.list_offsets.data_ptr(), stream))
        self._grouped = True

    # ----------------------------------------------------------------- search
    def probe(self, q, nprobe: Optional[int] = None) -> torch.Tensor:
        """The nprobe best lists of every query (int32 [nq, nprobe], on the GPU)."""
        nprobe = min(int(nprobe or self.nprobe), self.nlist)
        _, I = self.quantizer.search(self._to_device(q), nprobe)
        return I.to(torch.int32).contiguous()

    def search(self, q, k: int, nprobe: Optional[int] = None):
        host_out = not (isinstance(q, torch.Tensor) and q.is_cuda)
        as_numpy = isinstance(q, np.ndarray)
        qd = self._to_device(q)
        if self.ntotal == 0:
            # nothing to scan (an empty shard of a sharded index must still reach the collective): all padding
            nq = int(qd.shape[0]) if qd.dim() == 2 else 1
            D = torch.full((nq, k), float("inf") if self.metric == "l2" else float("-inf"), dtype=torch.float32,
                           device=self._dev())
            I = torch.full((nq, k), -1, dtype=torch.int64, device=self._dev())
            if host_out:
                D, I = D.cpu(), I.cpu()
                if as_numpy:
                    return D.numpy(), I.numpy()
            return D, I
        if not self._grouped:
            self._group()
        probes = self.probe(qd, nprobe)
        b = _Buf(qd, self.d, "q")
        D = torch.empty((b.n, k), dtype=torch.float32, device=self._dev())
        I = torch.empty((b.n, k), dtype=torch.int64, device=self._dev())
        stream = int(torch.cuda.current_stream(self.device).cuda_stream)
        _C.check(_C.lib().cvdb_index_search_lists(self.lists._h, b.ptr, b.n, b.dtype, int(k), probes.data_ptr(),
                                                  int(probes.shape[1]), D.data_ptr(), I.data_ptr(), 1, stream))
        if host_out:
            D, I = D.cpu(), I.cpu()
            if as_numpy:
                return D.numpy(), I.numpy()
        return D, I

    def list_of_row(self) -> torch.Tensor:
        """List id of every row, indexed by row id (int32, on the GPU)."""
        return self._assign

    def close(self) -> None:
        self.quantizer.close()
        self.lists.close()
