"""FAISS-style flat index surface over the C ABI (include/cvdb_b200.h).

Mirrors the convention BASELINE.json's north_star fixes for the hot path
(reference: README.md:2 "building the vectordb" - no reference API exists):
``IndexFlatIP(d)`` / ``IndexFlatL2(d)``, ``add(x)``, ``search(q, k) -> (D, I)``,
``ntotal``, ``d``, ``reset()``.

Inputs may be numpy arrays (host buffers: the C ABI does the host<->device
copies) or torch tensors (cuda tensors are used in place, on the current torch
stream).  float32 and bfloat16 are accepted.  Outputs come back in the same
container kind and place as the queries.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Tuple

import numpy as np

from . import _C

try:  # torch is plumbing (device memory, streams); numpy-only use works without it
    import torch
except Exception:  # pragma: no cover
    torch = None


def _is_torch(x) -> bool:
    return torch is not None and isinstance(x, torch.Tensor)


class _Buf:
    """A [n, d] row-major buffer the C ABI can read: pointer + dtype + place."""

    def __init__(self, x, d: Optional[int] = None, what: str = "x"):
        if _is_torch(x):
            if x.dim() == 1 and d is not None:
                x = x.view(1, -1)
            if x.dim() != 2:
                raise ValueError(f"{what} must be 2-D [n, d]")
            if x.dtype == torch.float32:
                self.dtype = _C.DTYPE_F32
            elif x.dtype == torch.bfloat16:
                self.dtype = _C.DTYPE_BF16
            elif x.dtype == torch.float16:
                self.dtype = _C.DTYPE_F16
            else:
                raise TypeError(f"{what}: dtype {x.dtype} not supported (float32, bfloat16 or float16)")
            x = x.contiguous()
            self.on_device = x.is_cuda
            self.device_index = x.device.index if x.is_cuda else None
            self.ptr = x.data_ptr()
            self.kind = "torch"
        else:
            x = np.asarray(x)
            if x.ndim == 1 and d is not None:
                x = x.reshape(1, -1)
            if x.ndim != 2:
                raise ValueError(f"{what} must be 2-D [n, d]")
            if x.dtype == np.float16:
                self.dtype = _C.DTYPE_F16
            else:
                if x.dtype != np.float32:
                    x = x.astype(np.float32)
                self.dtype = _C.DTYPE_F32
            x = np.ascontiguousarray(x)
            self.on_device = False
            self.device_index = None
            self.ptr = x.ctypes.data
            self.kind = "numpy"
        self.keep = x
        self.n, self.d = int(x.shape[0]), int(x.shape[1])
        if d is not None and self.d != d:
            raise ValueError(f"{what} has d={self.d}, index has d={d}")


def _i32_buf(v, n: int, like: _Buf, what: str):
    """int32 [n] vector living where `like` lives; returns (ptr, keepalive)."""
    if v is None:
        return None, None
    if like.on_device:
        t = torch.as_tensor(v, dtype=torch.int32, device=f"cuda:{like.device_index}").contiguous()
        if t.numel() != n:
            raise ValueError(f"{what} must have {n} entries")
        return t.data_ptr(), t
    if _is_torch(v):
        v = v.detach().cpu().numpy()
    a = np.ascontiguousarray(np.asarray(v, dtype=np.int32))
    if a.size != n:
        raise ValueError(f"{what} must have {n} entries")
    return a.ctypes.data, a


def _stream_for(buf: _Buf) -> int:
    if buf.on_device:
        return int(torch.cuda.current_stream(buf.device_index).cuda_stream)
    return 0


class IndexFlat:
    """Exact (brute-force) index.  ``storage="bf16"`` keeps rows as bf16 and
    scores with bf16 x bf16 -> fp32 tensor-core MMAs; ``storage="exact"`` keeps
    a three-way bf16 split of every fp32 value and reproduces fp32 scores."""

    def __init__(self, d: int, metric: str = "ip", storage: str = "bf16", device: int = 0):
        self._h = C.c_void_p()
        metric_code = {"ip": _C.METRIC_IP, "l2": _C.METRIC_L2}[metric.lower()]
        storage_code = {"bf16": _C.STORE_BF16, "exact": _C.STORE_EXACT}[storage.lower()]
        self.metric = metric.lower()
        self.storage = storage.lower()
        self.device = int(device)
        self._d = int(d)
        _C.check(_C.lib().cvdb_index_create(int(d), metric_code, storage_code, int(device), C.byref(self._h)))

    # -- FAISS-like members ------------------------------------------------
    @property
    def d(self) -> int:
        return self._d

    @property
    def ntotal(self) -> int:
        return int(_C.lib().cvdb_index_ntotal(self._h))

    def reset(self) -> None:
        _C.check(_C.lib().cvdb_index_reset(self._h))

    def reserve(self, n: int) -> None:
        _C.check(_C.lib().cvdb_index_reserve(self._h, int(n)))

    def truncate(self, n: int) -> None:
        """Drop the rows added last so that ``ntotal == n`` (keeps the allocation)."""
        _C.check(_C.lib().cvdb_index_truncate(self._h, int(n)))

    def nonfinite_rows(self, stream=None) -> int:
        """Rows seen so far (added or queried) that held a NaN / infinite element; waits for ``stream``."""
        out = C.c_int64(0)
        _C.check(_C.lib().cvdb_index_nonfinite_rows(self._h, C.byref(out), stream))
        return int(out.value)

    def add(self, x, check_finite: bool = False) -> None:
        """Append rows.  ``check_finite=True`` rejects the whole batch (ValueError, index unchanged) if any row
        holds a NaN or infinite value; the check is a counter kept by the packing kernel, it costs one
        stream synchronisation and no extra pass over the data."""
        b = _Buf(x, self._d, "x")
        self._check_place(b)
        st = _stream_for(b)
        if check_finite:
            n0, bad0 = self.ntotal, self.nonfinite_rows(st)
        _C.check(_C.lib().cvdb_index_add(self._h, b.ptr, b.n, b.dtype, int(b.on_device), st))
        if check_finite:
            bad = self.nonfinite_rows(st) - bad0
            if bad:
                self.truncate(n0)
                raise ValueError(f"add(): {bad} of {b.n} rows hold NaN or infinite values; nothing was added")

    def set_groups(self, group_db) -> None:
        """Group id per database row; search(..., group_q=) drops same-group rows."""
        if group_db is None:
            _C.check(_C.lib().cvdb_index_set_groups(self._h, None, 0, None))
            return
        if _is_torch(group_db) and group_db.is_cuda:
            t = group_db.to(torch.int32).contiguous()
            if t.numel() != self.ntotal:
                raise ValueError("group_db must have ntotal entries")
            _C.check(_C.lib().cvdb_index_set_groups(self._h, t.data_ptr(), 1,
                                                    int(torch.cuda.current_stream(self.device).cuda_stream)))
            torch.cuda.current_stream(self.device).synchronize()
            return
        if _is_torch(group_db):
            group_db = group_db.numpy()
        a = np.ascontiguousarray(np.asarray(group_db, dtype=np.int32))
        if a.size != self.ntotal:
            raise ValueError("group_db must have ntotal entries")
        _C.check(_C.lib().cvdb_index_set_groups(self._h, a.ctypes.data, 0, None))

    def search(self, q, k: int, *, self_ids=None, group_q=None, id_base: int = 0, profile: bool = False,
               force_slices: int = 0, force_variant: int = 0, debug_flags: int = 0,
               check_finite: bool = False) -> Tuple[object, object]:
        b = _Buf(q, self._d, "q")
        self._check_place(b)
        k = int(k)
        bad0 = self.nonfinite_rows(_stream_for(b)) if check_finite else 0
        opts = _C.SearchOpts()
        sp, keep_s = _i32_buf(self_ids, b.n, b, "self_ids")
        gp, keep_g = _i32_buf(group_q, b.n, b, "group_q")
        opts.self_ids = sp
        opts.group_q = gp
        opts.id_base = int(id_base)
        opts.profile = int(profile)
        opts.force_slices = int(force_slices)
        opts.force_variant = int(force_variant)
        opts.debug_flags = int(debug_flags)
        if b.on_device:
            dev = f"cuda:{b.device_index}"
            D = torch.empty((b.n, k), dtype=torch.float32, device=dev)
            I = torch.empty((b.n, k), dtype=torch.int64, device=dev)
            dp, ip = D.data_ptr(), I.data_ptr()
        else:
            Dn = np.empty((b.n, k), np.float32)
            In = np.empty((b.n, k), np.int64)
            dp, ip = Dn.ctypes.data, In.ctypes.data
        _C.check(_C.lib().cvdb_index_search(self._h, b.ptr, b.n, b.dtype, k, dp, ip, int(b.on_device), C.byref(opts),
                                            _stream_for(b)))
        del keep_s, keep_g
        if check_finite:
            bad = self.nonfinite_rows(_stream_for(b)) - bad0
            if bad:
                raise ValueError(f"search(): {bad} of {b.n} queries hold NaN or infinite values")
        if b.on_device:
            return D, I
        if b.kind == "torch":
            return torch.from_numpy(Dn), torch.from_numpy(In)
        return Dn, In

    # -- shard exchange in the engine's own candidate keys (include/cvdb_b200.h: cvdb_index_search_keys) ------
    def search_keys(self, q, k: int, *, self_ids=None, group_q=None, id_base: int = 0, profile: bool = False):
        """search() whose result stays in the kernel's 64-bit keys: int64 tensor [nq, k] on the GPU (the bit
        pattern of (ordered score << 32 | ~id), sorted best first, 0 = no result).  Ids include ``id_base``.
        This is what the ranks of a ShardedIndex exchange: one 8-byte word per candidate."""
        b = _Buf(q, self._d, "q")
        if not b.on_device:
            raise ValueError("search_keys() takes CUDA tensors (the exchange runs on the device)")
        self._check_place(b)
        k = int(k)
        opts = _C.SearchOpts()
        sp, keep_s = _i32_buf(self_ids, b.n, b, "self_ids")
        gp, keep_g = _i32_buf(group_q, b.n, b, "group_q")
        opts.self_ids, opts.group_q, opts.id_base, opts.profile = sp, gp, int(id_base), int(profile)
        keys = torch.empty((b.n, k), dtype=torch.int64, device=f"cuda:{b.device_index}")
        _C.check(_C.lib().cvdb_index_search_keys(self._h, b.ptr, b.n, b.dtype, k, keys.data_ptr(), C.byref(opts),
                                                 _stream_for(b)))
        del keep_s, keep_g
        return keys

    def merge_keys(self, keys, k: int):
        """k-way merge of key lists [nlists, nq, k_in] (an all_gather_into_tensor of every rank's search_keys
        result for the SAME queries) -> (D [nq, k] f32, I [nq, k] i64) on the GPU."""
        if not (_is_torch(keys) and keys.is_cuda and keys.dtype == torch.int64 and keys.dim() == 3):
            raise ValueError("keys must be a CUDA int64 tensor [nlists, nq, k_in]")
        keys = keys.contiguous()
        nl, nq, k_in = (int(v) for v in keys.shape)
        D = torch.empty((nq, int(k)), dtype=torch.float32, device=keys.device)
        I = torch.empty((nq, int(k)), dtype=torch.int64, device=keys.device)
        _C.check(_C.lib().cvdb_index_merge_keys(self._h, keys.data_ptr(), nq, nl, k_in, int(k), D.data_ptr(), I.data_ptr(),
                                                int(torch.cuda.current_stream(keys.device.index).cuda_stream)))
        return D, I

    def assign(self, x, return_dist: bool = True):
        """Nearest index row of every x (k = 1): (assign int32 [n], dist f32 [n])."""
        b = _Buf(x, self._d, "x")
        self._check_place(b)
        if b.on_device:
            dev = f"cuda:{b.device_index}"
            A = torch.empty((b.n,), dtype=torch.int32, device=dev)
            Dd = torch.empty((b.n,), dtype=torch.float32, device=dev) if return_dist else None
            ap, dp = A.data_ptr(), (Dd.data_ptr() if return_dist else None)
        else:
            An = np.empty((b.n,), np.int32)
            Dn = np.empty((b.n,), np.float32) if return_dist else None
            ap, dp = An.ctypes.data, (Dn.ctypes.data if return_dist else None)
        _C.check(_C.lib().cvdb_index_assign(self._h, b.ptr, b.n, b.dtype, ap, dp, int(b.on_device), _stream_for(b)))
        if b.on_device:
            return A, Dd
        if b.kind == "torch":
            return torch.from_numpy(An), (torch.from_numpy(Dn) if return_dist else None)
        return An, Dn

    # -- persistence -----------------------------------
    def save(self, path: str, chunk_rows: int = 1 << 20) -> None:
        """Dump the packed rows (exactly as stored in HBM) plus a small JSON header."""
        import json
        rb = int(_C.lib().cvdb_index_row_bytes(self._h))
        n = self.ntotal
        header = json.dumps({"format": "cvdb_b200.flat.v1", "d": self._d, "metric": self.metric,
                             "storage": self.storage, "ntotal": n, "row_bytes": rb}).encode()
        with open(path, "wb") as f:
            f.write(len(header).to_bytes(8, "little"))
            f.write(header)
            buf = np.empty((min(chunk_rows, max(n, 1)), rb), np.uint8)
            for r0 in range(0, n, chunk_rows):
                m = min(chunk_rows, n - r0)
                _C.check(_C.lib().cvdb_index_export_rows(self._h, r0, m, buf.ctypes.data, None))
                f.write(buf[:m].tobytes())

    @classmethod
    def load(cls, path: str, device: int = 0, chunk_rows: int = 1 << 20):
        import json
        with open(path, "rb") as f:
            hl = int.from_bytes(f.read(8), "little")
            h = json.loads(f.read(hl))
            if h.get("format") != "cvdb_b200.flat.v1":
                raise ValueError("not a cvdb_b200 index file")
            idx = IndexFlat(h["d"], h["metric"], h["storage"], device)
            rb, n = h["row_bytes"], h["ntotal"]
            if rb != int(_C.lib().cvdb_index_row_bytes(idx._h)):
                raise ValueError("row layout of the file does not match this build")
            idx.reserve(n)
            for r0 in range(0, n, chunk_rows):
                m = min(chunk_rows, n - r0)
                buf = np.frombuffer(f.read(m * rb), np.uint8)
                if buf.size != m * rb:
                    raise ValueError("truncated index file")
                _C.check(_C.lib().cvdb_index_import_rows(idx._h, buf.ctypes.data, m, None))
        return idx

    def add_from_file(self, path: str, dtype: str = "float32", chunk_rows: int = 1 << 20, offset: int = 0) -> int:
        """Streaming add() of a raw row-major [n, d] matrix on disk (float32 or bfloat16), memory-mapped
        and fed in chunks so the corpus never has to be resident on the host.  Returns the rows added."""
        esz = {"float32": 4, "bfloat16": 2, "float16": 2}[dtype]
        size = os.path.getsize(path) - offset
        if size % (esz * self._d):
            raise ValueError("file size is not a whole number of rows")
        n = size // (esz * self._d)
        mm = np.memmap(path, dtype=np.float32 if esz == 4 else np.uint16, mode="r", offset=offset, shape=(n, self._d))
        code = {"float32": _C.DTYPE_F32, "bfloat16": _C.DTYPE_BF16, "float16": _C.DTYPE_F16}[dtype]
        self.reserve(self.ntotal + int(n))
        for r0 in range(0, n, chunk_rows):
            # a row range of the mapping is contiguous already: the library reads it in place (its host threads
            # copy 64 MB pieces into pinned staging buffers while the previous piece is on the PCIe bus)
            blk = mm[r0:r0 + chunk_rows]
            _C.check(_C.lib().cvdb_index_add(self._h, blk.ctypes.data, blk.shape[0], code, 0, None))
        return int(n)

    # -- measurement hooks ---------------------------------------------------
    def last_kernel_ms(self) -> float:
        return float(_C.lib().cvdb_index_last_kernel_ms(self._h))

    def profile_ms(self) -> list:
        """Device times (ms) of the profiled kernel launches since the previous call."""
        buf = (C.c_float * 64)()
        n = _C.lib().cvdb_index_profile_ms(self._h, buf, 64)
        if n < 0:
            _C.check(n)
        return [float(buf[i]) for i in range(n)]

    def last_work(self) -> dict:
        f, by, s, g = C.c_double(), C.c_double(), C.c_int(), C.c_int()
        _C.check(_C.lib().cvdb_index_last_work(self._h, C.byref(f), C.byref(by), C.byref(s), C.byref(g)))
        return {"flops": f.value, "db_bytes": by.value, "n_slices": s.value, "grid": g.value,
                "variant": int(_C.lib().cvdb_index_last_variant(self._h))}

    # -- internals -----------------------------------------------------------
    def _check_place(self, b: _Buf) -> None:
        if b.on_device and b.device_index != self.device:
            raise ValueError(f"tensor lives on cuda:{b.device_index}, index on cuda:{self.device}")

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            _C.lib().cvdb_index_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass


class IndexFlatIP(IndexFlat):
    def __init__(self, d: int, storage: str = "bf16", device: int = 0):
        super().__init__(d, "ip", storage, device)


class IndexFlatL2(IndexFlat):
    def __init__(self, d: int, storage: str = "bf16", device: int = 0):
        super().__init__(d, "l2", storage, device)


def merge_topk(D_parts, I_parts, k: int, metric: str = "ip"):
    """k-way select over per-shard results stacked as [nlists, nq, k_in]
    (numpy -> host path, torch cuda -> device path on the current stream)."""
    metric_code = {"ip": _C.METRIC_IP, "l2": _C.METRIC_L2}[metric.lower()]
    if _is_torch(D_parts):
        Dc = D_parts.to(torch.float32).contiguous()
        Ic = I_parts.to(torch.int64).contiguous()
        nl, nq, k_in = Dc.shape
        if Dc.is_cuda:
            D = torch.empty((nq, k), dtype=torch.float32, device=Dc.device)
            I = torch.empty((nq, k), dtype=torch.int64, device=Dc.device)
            _C.check(_C.lib().cvdb_merge_topk(Dc.data_ptr(), Ic.data_ptr(), nq, nl, k_in, int(k), metric_code,
                                              D.data_ptr(), I.data_ptr(), 1,
                                              int(torch.cuda.current_stream(Dc.device.index).cuda_stream)))
            return D, I
        D, I = merge_topk(Dc.numpy(), Ic.numpy(), k, metric)
        return torch.from_numpy(D), torch.from_numpy(I)
    Dc = np.ascontiguousarray(np.asarray(D_parts, np.float32))
    Ic = np.ascontiguousarray(np.asarray(I_parts, np.int64))
    nl, nq, k_in = Dc.shape
    D = np.empty((nq, k), np.float32)
    I = np.empty((nq, k), np.int64)
    _C.check(_C.lib().cvdb_merge_topk(Dc.ctypes.data, Ic.ctypes.data, nq, nl, k_in, int(k), metric_code, D.ctypes.data,
                                      I.ctypes.data, 0, None))
    return D, I
