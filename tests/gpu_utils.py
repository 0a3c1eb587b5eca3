"""Helpers shared by the -m gpu tests: build gdf_columns over torch CUDA tensors and call the C ABI
the way the reference's own python tests do (reference: libgdf/python/tests/utils.py)."""
import numpy as np
import torch

from libgdf_b200 import columns as C
from libgdf_b200.libgdf_cffi import ffi, libgdf

NP_TO_GDFNAME = {np.dtype(np.int8): "GDF_INT8", np.dtype(np.int16): "GDF_INT16", np.dtype(np.int32): "GDF_INT32",
                 np.dtype(np.int64): "GDF_INT64", np.dtype(np.float32): "GDF_FLOAT32", np.dtype(np.float64): "GDF_FLOAT64"}


def gen_rand(dtype, size, low=-10000, high=10000, rng=None):
    """Same value ranges as the reference's gen_rand (python/tests/utils.py:36-47)."""
    rng = rng or np.random
    dtype = np.dtype(dtype)
    if dtype.kind == "f":
        return rng.rand(size).astype(dtype) * 2 - 1
    if dtype == np.int8:
        low, high = max(low, -128), min(high, 128)
    return rng.randint(low=low, high=high, size=size).astype(dtype)


def rand_mask(n, p_valid=0.7, rng=None):
    rng = rng or np.random
    bits = rng.rand(n) < p_valid
    return np.packbits(bits, bitorder="little"), bits


def make_context(method="GDF_HASH", sort_result=0, distinct=0):
    ctx = ffi.new("gdf_context*")
    libgdf.gdf_context_view(ctx, 0, getattr(libgdf, method), distinct, sort_result, 0)
    return ctx


def join(kind, left_cols, right_cols, left_valid=None, right_valid=None, dtypes=None, api=None, method="GDF_HASH"):
    """Call gdf_{inner,left,full}_join on key columns only; returns (left_idx, right_idx) numpy."""
    api = api or libgdf
    nk = len(left_cols)
    left_valid = left_valid or [None] * nk
    right_valid = right_valid or [None] * nk
    dtypes = dtypes or [None] * nk
    L = [C.column(c, v, dtype=d, api=api) for c, v, d in zip(left_cols, left_valid, dtypes)]
    R = [C.column(c, v, dtype=d, api=api) for c, v, d in zip(right_cols, right_valid, dtypes)]
    la, ra = C.column_array(L), C.column_array(R)
    idx = ffi.new("int[]", list(range(nk)))
    out_l, out_r = ffi.new("gdf_column*"), ffi.new("gdf_column*")
    ctx = ffi.new("gdf_context*")
    api.gdf_context_view(ctx, 0, getattr(api, method), 0, 0, 0)
    fn = getattr(api, "gdf_%s_join" % kind)
    err = fn(la, nk, idx, ra, nk, idx, nk, 0, ffi.NULL, out_l, out_r, ctx)
    if err not in (None, 0):
        raise RuntimeError("gdf_%s_join -> %r" % (kind, err))
    assert out_l.size == out_r.size
    res = []
    for o in (out_l, out_r):
        n = int(o.size)
        if n == 0:
            res.append(np.empty(0, np.int32))
            if o.data != ffi.NULL:
                api.gdf_column_free(o)
            continue
        assert o.dtype == api.GDF_INT32 and o.valid == ffi.NULL and o.null_count == 0
        t = C.alias_column_data(o, np.int32).clone()
        torch.cuda.synchronize()
        res.append(t.cpu().numpy())
        api.gdf_column_free(o)
    return res[0], res[1]


def sorted_pairs(l, r):
    order = np.lexsort((r, l))
    return np.stack([l[order], r[order]], axis=1)


GROUPBY_FN = {"sum": "gdf_group_by_sum", "min": "gdf_group_by_min", "max": "gdf_group_by_max",
              "avg": "gdf_group_by_avg", "count": "gdf_group_by_count"}


def groupby(op, key_cols, values, out_np_dtype=None, api=None, key_dtypes=None, sort_result=0, method="GDF_HASH",
            distinct=0, want_indices=False):
    """Call gdf_group_by_<op>; returns (list of key arrays, agg array) trimmed to the group count, in the order
    the library wrote them (+ the out_col_indices array, size_t, with want_indices)."""
    api = api or libgdf
    n = len(values)
    key_dtypes = key_dtypes or [None] * len(key_cols)
    K = [C.column(k, dtype=d, api=api) for k, d in zip(key_cols, key_dtypes)]
    V = C.column(values, api=api)
    if out_np_dtype is None:
        out_np_dtype = values.dtype
    OK = [C.empty_column(n, K[i].data.dtype, dtype=key_dtypes[i], api=api) for i in range(len(K))]
    OA = C.empty_column(n, getattr(torch, np.dtype(out_np_dtype).name), api=api)
    OI = C.empty_column(n, torch.int64, api=api) if want_indices else None
    ctx = ffi.new("gdf_context*")
    api.gdf_context_view(ctx, 0, getattr(api, method), distinct, sort_result, 0)
    ka, oka = C.column_array(K), C.column_array(OK)
    err = getattr(api, GROUPBY_FN[op])(len(K), ka, V.cdata, OI.cdata if OI else ffi.NULL, oka, OA.cdata, ctx)
    if err not in (None, 0):
        raise RuntimeError("%s -> %r" % (GROUPBY_FN[op], err))
    torch.cuda.synchronize()
    g = int(OA.cdata.size)
    for o in OK:
        assert int(o.cdata.size) == g
    keys, agg = [o.data[:g].cpu().numpy() for o in OK], OA.data[:g].cpu().numpy()
    if want_indices:
        assert int(OI.cdata.size) == g
        return keys, agg, OI.data[:g].cpu().numpy()
    return keys, agg


def rows_as_sorted_tuples(key_arrays, agg):
    rows = list(zip(*[k.tolist() for k in key_arrays], agg.tolist()))
    return sorted(rows)
