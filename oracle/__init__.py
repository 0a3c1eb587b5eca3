"""TEST INFRASTRUCTURE ONLY - CPU oracle for the gdf hot path.

Nothing under ``libgdf_b200/`` imports this package.  Allowed users: ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.

Two halves:

* ``liboracle.so`` (``gdf_oracle.c``, plain C): row hashing, hash-partition ids, hash join, hash
  group-by - the loop-heavy algorithms, each function citing the reference file:line it restates;
* ``np_oracle`` (numpy): the element-wise / scan operators (binary ops, reductions, comparisons,
  gdf_filter, gpu_apply_stencil).

Parity status: pinned.  The oracle is checked against the reference's own golden vectors
(``tests/golden/reference_vectors.json``) and MurmurHash3 known answers in ``tests/test_oracle.py``
(CPU), and against the reference's own kernels rebuilt for sm_100a (``oracle/_ref/libgdf_ref.so``) in
``tests/test_reference_parity.py`` (GPU).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liboracle.so")

# gdf_dtype codes (include/gdf/cffi/types.h)
GDF_INT8, GDF_INT16, GDF_INT32, GDF_INT64, GDF_FLOAT32, GDF_FLOAT64, GDF_DATE32, GDF_DATE64, GDF_TIMESTAMP = range(1, 10)
NP_TO_GDF = {np.dtype(np.int8): GDF_INT8, np.dtype(np.int16): GDF_INT16, np.dtype(np.int32): GDF_INT32,
             np.dtype(np.int64): GDF_INT64, np.dtype(np.float32): GDF_FLOAT32, np.dtype(np.float64): GDF_FLOAT64}
GDF_TO_NP = {v: k for k, v in NP_TO_GDF.items()}
GDF_TO_NP.update({GDF_DATE32: np.dtype(np.int32), GDF_DATE64: np.dtype(np.int64), GDF_TIMESTAMP: np.dtype(np.int64)})

OP_SUM, OP_MIN, OP_MAX, OP_AVG, OP_COUNT = 0, 1, 2, 3, 4   # gdf_agg_op
JOIN_INNER, JOIN_LEFT, JOIN_FULL = 0, 1, 2


def build():
    """Compile liboracle.so if it is missing or older than its source."""
    src = os.path.join(_HERE, "gdf_oracle.c")
    if not os.path.isfile(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.orc_murmur3_32.restype = ctypes.c_uint32
        _lib.orc_join.restype = ctypes.c_longlong
        _lib.orc_groupby.restype = ctypes.c_longlong
        _lib.orc_filter_i64.restype = ctypes.c_size_t
        _lib.orc_sum_i64.restype = ctypes.c_int64
    return _lib


def _ptr_array(arrays):
    """(void*[]) over numpy arrays (None -> NULL); returns (ctypes array, keep-alive list)."""
    keep = [None if a is None else np.ascontiguousarray(a) for a in arrays]
    arr = (ctypes.c_void_p * len(keep))(*[None if a is None else a.ctypes.data for a in keep])
    return arr, keep


def _dtypes(cols, dtypes=None):
    if dtypes is None:
        dtypes = [NP_TO_GDF[np.asarray(c).dtype] for c in cols]
    return (ctypes.c_int * len(cols))(*dtypes)


def murmur3_32(data: bytes) -> int:
    return int(lib().orc_murmur3_32(ctypes.c_char_p(data), ctypes.c_int(len(data))))


def hash_rows(cols, identity=False, dtypes=None):
    """gdf_hash: int32 hash per row (reference: src/hashing.cu:83-154)."""
    n = len(cols[0])
    out = np.empty(n, dtype=np.int32)
    ptrs, keep = _ptr_array(cols)
    lib().orc_hash_rows(len(cols), ptrs, _dtypes(keep, dtypes), ctypes.c_size_t(n), int(identity),
                        out.ctypes.data_as(ctypes.c_void_p))
    return out


def partition_ids(cols, num_partitions, identity=False, dtypes=None):
    """Partition number of every row under gdf_hash_partition (reference: src/hashing.cu:196-320)."""
    n = len(cols[0])
    out = np.empty(n, dtype=np.int32)
    ptrs, keep = _ptr_array(cols)
    lib().orc_partition_ids(len(cols), ptrs, _dtypes(keep, dtypes), ctypes.c_size_t(n), int(identity),
                            int(num_partitions), out.ctypes.data_as(ctypes.c_void_p))
    return out


def join(kind, left_cols, right_cols, left_valid=None, right_valid=None, dtypes=None):
    """Hash join on the given key columns -> (left_idx, right_idx) int32 arrays, unspecified order
    (reference: src/join/hash/join_kernels.cuh:48-455, join_compute_api.h:147-186)."""
    nl, nr = len(left_cols[0]), len(right_cols[0])
    lp, lkeep = _ptr_array(left_cols)
    rp, rkeep = _ptr_array(right_cols)
    lvp = rvp = None
    if left_valid is not None and any(v is not None for v in left_valid):
        lvp, lvkeep = _ptr_array(left_valid)
    if right_valid is not None and any(v is not None for v in right_valid):
        rvp, rvkeep = _ptr_array(right_valid)
    cap = max(nl + nr, 16)
    while True:
        ol = np.empty(cap, dtype=np.int32)
        orr = np.empty(cap, dtype=np.int32)
        cnt = lib().orc_join(int(kind), len(left_cols), _dtypes(lkeep, dtypes), lp, lvp, ctypes.c_size_t(nl),
                             rp, rvp, ctypes.c_size_t(nr), ol.ctypes.data_as(ctypes.c_void_p),
                             orr.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(cap))
        if cnt == -1:
            cap *= 4
            continue
        if cnt < 0:
            raise MemoryError("oracle join allocation failed")
        return ol[:cnt].copy(), orr[:cnt].copy()


def groupby(op, key_cols, values, out_dtype=None, key_dtypes=None, val_dtype=None):
    """Hash group-by with one aggregate -> (list of key arrays, aggregate array), first-seen order
    (reference: src/groupby/hash/groupby_kernels.cuh:47-160, groupby/groupby.cuh:88-190,308-386)."""
    n = len(key_cols[0])
    values = np.ascontiguousarray(values)
    val_dtype = NP_TO_GDF[values.dtype] if val_dtype is None else val_dtype
    if out_dtype is None:
        out_dtype = val_dtype
    kp, kkeep = _ptr_array(key_cols)
    out_keys = [np.empty(n, dtype=a.dtype) for a in kkeep]
    okp, _ = _ptr_array(out_keys)
    # _ptr_array copies if not contiguous; out_keys are fresh so pointers are stable
    okp = (ctypes.c_void_p * len(out_keys))(*[a.ctypes.data for a in out_keys])
    agg_np = GDF_TO_NP[out_dtype if op in (OP_COUNT, OP_AVG) else val_dtype]
    out_agg = np.empty(n, dtype=agg_np)
    g = lib().orc_groupby(int(op), len(key_cols), _dtypes(kkeep, key_dtypes), kp, ctypes.c_size_t(n),
                          int(val_dtype), values.ctypes.data_as(ctypes.c_void_p), int(out_dtype), okp,
                          out_agg.ctypes.data_as(ctypes.c_void_p))
    if g < 0:
        raise MemoryError("oracle groupby allocation failed")
    return [k[:g].copy() for k in out_keys], out_agg[:g].copy()


def filter_i64(data, value):
    data = np.ascontiguousarray(data, dtype=np.int64)
    out = np.empty(len(data), dtype=np.uint64)
    k = lib().orc_filter_i64(data.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(len(data)),
                             ctypes.c_int64(int(value)), out.ctypes.data_as(ctypes.c_void_p))
    return out[:k].copy()


def sum_i64(data, valid=None):
    data = np.ascontiguousarray(data, dtype=np.int64)
    vp = None if valid is None else np.ascontiguousarray(valid).ctypes.data_as(ctypes.c_void_p)
    return int(lib().orc_sum_i64(data.ctypes.data_as(ctypes.c_void_p), vp, ctypes.c_size_t(len(data))))
