"""ctypes binding of libcvdb_b200.so (include/cvdb_b200.h).  No logic lives here.

The library is built in-tree by ``__graft_entry__.build()`` (or
``make -C cloudvectordb_b200/csrc``).  There is no CPU fallback: if the shared
library is missing, or no B200 is present, the product path raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CVDB_LIB_PATH") or os.path.join(_HERE, "libcvdb_b200.so")  # env override: A/B of builds

OK, EINVAL, ECUDA, ENOMEM, ELIMIT = 0, -1, -2, -3, -4
METRIC_IP, METRIC_L2 = 0, 1
DTYPE_F32, DTYPE_BF16, DTYPE_F16 = 0, 1, 2
STORE_BF16, STORE_EXACT = 0, 1
MAX_K = 2048


class CvdbError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"cvdb error {code}: {msg}")
        self.code = code


class SearchOpts(C.Structure):
    _fields_ = [
        ("self_ids", C.c_void_p),
        ("group_q", C.c_void_p),
        ("id_base", C.c_int64),
        ("profile", C.c_int),
        ("force_slices", C.c_int),
        ("force_variant", C.c_int),
        ("debug_flags", C.c_int),
    ]


# every symbol include/cvdb_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "cvdb_index_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "cvdb_index_destroy": (C.c_int, [C.c_void_p]),
    "cvdb_index_reset": (C.c_int, [C.c_void_p]),
    "cvdb_index_reserve": (C.c_int, [C.c_void_p, C.c_int64]),
    "cvdb_index_truncate": (C.c_int, [C.c_void_p, C.c_int64]),
    "cvdb_index_nonfinite_rows": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.c_void_p]),
    "cvdb_index_ntotal": (C.c_int64, [C.c_void_p]),
    "cvdb_index_dim": (C.c_int, [C.c_void_p]),
    "cvdb_index_add": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p]),
    "cvdb_index_set_groups": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "cvdb_index_search": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                    C.c_int, C.POINTER(SearchOpts), C.c_void_p]),
    "cvdb_index_assign": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                    C.c_void_p]),
    "cvdb_index_last_kernel_ms": (C.c_float, [C.c_void_p]),
    "cvdb_index_profile_ms": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.c_int]),
    "cvdb_index_last_work": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int),
                                       C.POINTER(C.c_int)]),
    "cvdb_index_last_variant": (C.c_int, [C.c_void_p]),
    "cvdb_index_group_by_list": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "cvdb_index_list_offsets": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "cvdb_index_search_lists": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                          C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "cvdb_build_triplets": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int64, C.c_int, C.c_int,
                                      C.c_int, C.c_float, C.c_int, C.c_void_p, C.c_void_p]),
    "cvdb_index_row_bytes": (C.c_int64, [C.c_void_p]),
    "cvdb_index_export_rows": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]),
    "cvdb_index_import_rows": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "cvdb_plan_slices": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int64, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "cvdb_merge_topk": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                  C.c_void_p, C.c_int, C.c_void_p]),
    "cvdb_index_search_keys": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p,
                                         C.POINTER(SearchOpts), C.c_void_p]),
    "cvdb_index_merge_keys": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                        C.c_void_p, C.c_void_p]),
    "cvdb_selfjoin_begin": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "cvdb_selfjoin_chunk": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]),
    "cvdb_selfjoin_seed": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]),
    "cvdb_selfjoin_cross": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64,
                                      C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]),
    "cvdb_merge_keys": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "cvdb_selfjoin_finish": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cvdb_selfjoin_dirty": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int64), C.c_void_p]),
    "cvdb_selfjoin_end": (C.c_int, [C.c_void_p]),
    "cvdb_kmeans_accumulate": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_void_p]),
    "cvdb_kmeans_finalize": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "cvdb_kmeans_split_empty": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p]),
    "cvdb_last_error": (C.c_char_p, []),
    "cvdb_kernel_launches": (C.c_int64, []),
    "cvdb_version": (C.c_int, []),
}

_lib = None


def lib() -> C.CDLL:
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C cloudvectordb_b200/csrc`.  There is no CPU fallback.")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(code: int) -> None:
    if code != OK:
        raise CvdbError(code, lib().cvdb_last_error().decode("utf-8", "replace"))
