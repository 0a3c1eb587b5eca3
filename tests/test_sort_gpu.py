"""-m gpu: row ordering and concatenation (SURVEY.md 8f ranks 2 and 3, and the flag_sort_result / sorted-AVG part of
row a9) through the C ABI, against numpy, the CPU oracle, the reference's golden literals and - where the reference
fixes the answer - the reference's own kernels (oracle/_ref/libgdf_ref.so).

  gdf_order_by                 sqls_ops.cu:1373-1392, sqls_rtti_comp.hpp:299-320 (ties unspecified in the reference)
  flag_sort_result / AVG       groupby_compute_api.h:211-222, groupby.cuh:346-386 (output in lexicographic key order)
  GDF_SORT group-by            sqls_ops.cu:1134-1289; golden literals sqls_g_tester.cu:114-263
  gdf_column_concat / mask     column.cpp:53-153, validops.cu:203-256
"""
import json
import os

import numpy as np
import pytest
import torch

import oracle
from libgdf_b200 import columns as C
from libgdf_b200.libgdf_cffi import GDFError, ffi, libgdf
import gpu_utils as G
from test_reference_parity import ref  # noqa: F401  (fixture: the reference build)

pytestmark = pytest.mark.gpu
GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.json")))
OPS = {"sum": oracle.OP_SUM, "min": oracle.OP_MIN, "max": oracle.OP_MAX, "avg": oracle.OP_AVG, "count": oracle.OP_COUNT}


def order_by(cols_np, api=None):
    api = api or libgdf
    n = len(cols_np[0])
    cols = [C.column(c, api=api) for c in cols_np]
    d_cols = torch.zeros(len(cols), dtype=torch.int64, device="cuda")
    d_types = torch.zeros(len(cols), dtype=torch.int32, device="cuda")
    d_indx = torch.full((max(n, 1),), -1, dtype=torch.int64, device="cuda")
    rc = api.gdf_order_by(n, C.struct_array(cols), len(cols), ffi.cast("void**", d_cols.data_ptr()),
                          ffi.cast("int*", d_types.data_ptr()), ffi.cast("size_t*", d_indx.data_ptr()))
    assert rc in (None, 0)
    torch.cuda.synchronize()
    assert d_cols.cpu().tolist() == [c.data.data_ptr() for c in cols]        # soa_col_info side effect
    assert d_types.cpu().tolist() == [int(c.cdata.dtype) for c in cols]
    return d_indx[:n].cpu().numpy()


@pytest.mark.parametrize("types", [[np.int64], [np.int32], [np.int8], [np.int16], [np.float32], [np.float64],
                                   [np.int32, np.int64], [np.int8, np.float64, np.int16], [np.int64, np.int32, np.float32]],
                         ids=lambda ts: "-".join(np.dtype(t).name for t in ts))
@pytest.mark.parametrize("n", [1, 31, 2049, 100_003])
def test_order_by_matches_stable_lexsort(types, n):
    cols = [G.gen_rand(t, n, -50, 50) if np.dtype(t).kind != "f" else np.round(G.gen_rand(t, n) * 20) / 4 + 0.0 for t in types]   # + 0.0: no -0.0 (a tie under `<`, ordered first here)
    got = order_by(cols)
    want = np.lexsort(tuple(reversed(cols)))                 # stable, first column most significant
    np.testing.assert_array_equal(got, want)


def test_order_by_wide_values_and_extremes():
    n = 50_000
    a = np.random.randint(-2 ** 62, 2 ** 62, n).astype(np.int64)
    a[:4] = [np.iinfo(np.int64).min, np.iinfo(np.int64).max, 0, -1]
    f = (np.random.rand(n) - 0.5) * 1e300
    f[:5] = [-np.inf, np.inf, -0.0, 0.0, 1e-310]
    np.testing.assert_array_equal(order_by([a]), np.argsort(a, kind="stable"))
    got = order_by([f])
    assert (np.diff(f[got]) >= 0).all() and sorted(got.tolist()) == list(range(n))
    assert list(f[got][:1]) == [-np.inf] and f[got][-1] == np.inf
    zeros = [i for i in got if f[i] == 0]
    assert np.signbit(f[zeros[0]]) and not np.signbit(f[zeros[1]])           # -0.0 before +0.0 (total order)


def test_order_by_against_the_reference_build(ref):  # noqa: F811
    n = 30_011
    cols = [np.random.randint(0, 40, n).astype(np.int32), np.random.randint(-5, 5, n).astype(np.int64),
            np.round(np.random.rand(n) * 8) / 8]
    mine, theirs = order_by(cols), order_by(cols, api=ref)
    assert sorted(theirs.tolist()) == list(range(n))
    for c in cols:                                            # same sequence of keys; ties are unordered in the reference
        np.testing.assert_array_equal(c[mine], c[theirs])


def test_order_by_rejects_masks_and_handles_empty():
    col = C.column(np.arange(8, dtype=np.int32), np.array([0xff], np.uint8))
    idx = torch.zeros(8, dtype=torch.int64, device="cuda")
    with pytest.raises(GDFError) as e:
        libgdf.gdf_order_by(8, C.struct_array([col]), 1, ffi.NULL, ffi.NULL, ffi.cast("size_t*", idx.data_ptr()))
    assert e.value.errcode == "GDF_VALIDITY_UNSUPPORTED"
    assert len(order_by([np.empty(0, np.int64)])) == 0


# ---- hash group-by, flag_sort_result = 1 and AVG: ORDERED output ----
def _ordered_rows(keys, agg):
    return list(zip(*[k.tolist() for k in keys], agg.tolist()))


@pytest.mark.parametrize("key_types", [[np.int64], [np.int32, np.int64], [np.int8, np.float64, np.int32]],
                         ids=lambda ts: "-".join(np.dtype(t).name for t in ts))
@pytest.mark.parametrize("op", ["sum", "avg", "min", "count"])
def test_hash_groupby_sort_result_is_ordered_like_the_reference(ref, key_types, op):  # noqa: F811
    n = 60_000
    keys = [np.random.randint(-20, 20 if len(key_types) > 1 else 3000, n).astype(t) for t in key_types]
    vals = G.gen_rand(np.int64, n, 0, 1000)
    out_t = np.float64 if op == "avg" else (np.int32 if op == "count" else np.int64)
    gk, ga = G.groupby(op, keys, vals, out_t, sort_result=1)
    rk, ra = G.groupby(op, keys, vals, out_t, sort_result=1, api=ref)
    got, want = _ordered_rows(gk, ga), _ordered_rows(rk, ra)
    assert [r[:-1] for r in got] == sorted(r[:-1] for r in got), "not in lexicographic key order"
    if op == "avg":
        assert [r[:-1] for r in got] == [r[:-1] for r in want]
        np.testing.assert_allclose([r[-1] for r in got], [r[-1] for r in want], rtol=1e-9)
    else:
        assert got == want                                    # ordered, bit-exact


def test_hash_groupby_avg_is_sorted_without_the_flag(ref):  # noqa: F811
    n = 20_000
    keys = [np.random.permutation(n).astype(np.int64) % 977]
    vals = G.gen_rand(np.float64, n)
    gk, ga = G.groupby("avg", keys, vals, np.float64, sort_result=0)
    rk, ra = G.groupby("avg", keys, vals, np.float64, sort_result=0, api=ref)
    np.testing.assert_array_equal(gk[0], np.sort(gk[0]))
    np.testing.assert_array_equal(gk[0], rk[0])
    np.testing.assert_allclose(ga, ra, rtol=1e-9)


def test_sort_result_on_the_fast_path_at_scale():
    """single int64 key, int64 values (the C4 kernel), 3e6 rows / 2e5 groups, ordered output vs numpy."""
    n, groups = 3_000_000, 200_000
    keys = (np.random.randint(0, groups, n).astype(np.int64) - groups // 2) * 1_000_003
    vals = np.random.randint(0, 1000, n).astype(np.int64)
    gk, ga = G.groupby("sum", [keys], vals, sort_result=1)
    uk, inv = np.unique(keys, return_inverse=True)
    np.testing.assert_array_equal(gk[0], uk)
    np.testing.assert_array_equal(ga, np.bincount(inv, weights=None, minlength=len(uk)) * 0 + np.bincount(inv, vals).astype(np.int64))


# ---- GDF_SORT method ----
@pytest.mark.parametrize("op", ["sum", "min", "max", "count", "avg"])
def test_sort_method_reference_golden_literals(op):
    g = GOLDEN["groupby"]
    keys = [np.array(c, dtype=t) for c, t in zip(g["keys"], g["key_dtypes"])]
    vals = np.array(g["agg_min_max" if op in ("min", "max") else "agg_sum_count_avg"], dtype=np.float64)
    out_t = np.int32 if op == "count" else np.float64
    gk, ga, gi = G.groupby(op, keys, vals, out_t, method="GDF_SORT", want_indices=True)
    assert _ordered_rows(gk, ga) == list(zip(*g["group_keys_sorted"], g[op]))      # in the reference's (sorted) order
    for row, i in enumerate(gi):                               # one row of each group (the reference: rows 5,0,2,4)
        assert all(k[int(i)] == gk[c][row] for c, k in enumerate(keys))


@pytest.mark.parametrize("op", ["sum", "min", "max", "count"])
@pytest.mark.parametrize("val_t", [np.int32, np.int64, np.float64])
def test_sort_method_against_the_reference_build(ref, op, val_t):  # noqa: F811
    n = 30_000
    keys = [np.random.randint(0, 12, n).astype(np.int32), np.random.randint(-3, 3, n).astype(np.int64)]
    vals = G.gen_rand(val_t, n, 0, 1000)
    out_t = np.int32 if op == "count" else val_t
    gk, ga, gi = G.groupby(op, keys, vals, out_t, method="GDF_SORT", want_indices=True)
    rk, ra, ri = G.groupby(op, keys, vals, out_t, method="GDF_SORT", want_indices=True, api=ref)
    for a, b in zip(gk, rk):
        np.testing.assert_array_equal(a, b)
    if np.dtype(val_t).kind == "f" and op == "sum":
        np.testing.assert_allclose(ga, ra, rtol=1e-9)
    else:
        np.testing.assert_array_equal(ga, ra)
    for row in range(len(ga)):                                 # both name a row of the right group
        for c, k in enumerate(keys):
            assert k[int(gi[row])] == gk[c][row] == k[int(ri[row])]


def test_sort_method_count_distinct(ref):  # noqa: F811
    n = 10_000
    keys = [np.random.randint(0, 123, n).astype(np.int32)]
    vals = G.gen_rand(np.int64, n, 0, 10)
    _, ga = G.groupby("count", keys, vals, np.int32, method="GDF_SORT", distinct=1)
    _, ra = G.groupby("count", keys, vals, np.int32, method="GDF_SORT", distinct=1, api=ref)
    assert len(ga) == 1 == len(ra) and int(ga[0]) == int(ra[0]) == len(np.unique(keys[0]))


# ---- concatenation ----
def _concat(api, parts, masks, out_mask=True):
    cols = [C.column(p, m, api=api) for p, m in zip(parts, masks)]
    total = sum(len(p) for p in parts)
    out = C.empty_column(total, cols[0].data.dtype, with_valid=out_mask, api=api)
    if out_mask:
        out.valid.fill_(0x5a)                                  # stale bits must be overwritten
    rc = api.gdf_column_concat(out.cdata, C.column_array(cols), len(cols))
    assert rc in (None, 0)
    torch.cuda.synchronize()
    valid = np.unpackbits(out.valid.cpu().numpy(), bitorder="little")[:total] if out_mask else None
    return out.data.cpu().numpy(), valid, int(out.cdata.null_count), (out.valid.cpu().numpy() if out_mask else None)


@pytest.mark.parametrize("np_t", [np.int8, np.int32, np.int64, np.float64])
@pytest.mark.parametrize("sizes", [[5], [8, 8], [3, 0, 70, 1], [1000, 17, 4099, 64, 1]])
def test_column_concat_with_masks(ref, np_t, sizes):  # noqa: F811
    parts = [G.gen_rand(np_t, n, 0, 100) for n in sizes]
    masks, bits = [], []
    for k, n in enumerate(sizes):
        if k % 3 == 1:                                         # a column WITHOUT mask counts as all valid
            masks.append(None)
            bits.append(np.ones(n, bool))
        else:
            m, b = G.rand_mask(n, 0.6)
            masks.append(m)
            bits.append(b)
    if all(m is None for m in masks):
        masks[0], bits[0] = G.rand_mask(sizes[0], 0.6)
    data, valid, nulls, raw = _concat(libgdf, parts, masks)
    np.testing.assert_array_equal(data, np.concatenate(parts))
    np.testing.assert_array_equal(valid.astype(bool), np.concatenate(bits))
    assert nulls == sum(int((~b).sum()) for m, b in zip(masks, bits) if m is not None)
    total = sum(sizes)
    if total % 8:
        assert raw[-1] >> (total % 8) == 0                     # bits past the end are zero (validops.cu:218)
    rdata, rvalid, rnulls, _ = _concat(ref, parts, masks)
    np.testing.assert_array_equal(data, rdata)
    np.testing.assert_array_equal(valid, rvalid)
    assert nulls == rnulls


def test_column_concat_without_masks_fills_the_output_mask():
    parts = [np.arange(10, dtype=np.int32), np.arange(7, dtype=np.int32)]
    data, valid, nulls, _ = _concat(libgdf, parts, [None, None])
    np.testing.assert_array_equal(data, np.concatenate(parts))
    assert valid.all() and nulls == 0


def test_column_concat_error_codes():
    a, b = C.column(np.zeros(4, np.int32)), C.column(np.zeros(4, np.int64))
    out = C.empty_column(8, torch.int32)
    with pytest.raises(GDFError) as e:
        libgdf.gdf_column_concat(out.cdata, C.column_array([a, b]), 2)
    assert e.value.errcode == "GDF_DTYPE_MISMATCH"
    small = C.empty_column(7, torch.int32)
    with pytest.raises(GDFError) as e:
        libgdf.gdf_column_concat(small.cdata, C.column_array([a, a]), 2)
    assert e.value.errcode == "GDF_COLUMN_SIZE_MISMATCH"
    with pytest.raises(GDFError) as e:
        libgdf.gdf_column_concat(out.cdata, ffi.NULL, 2)
    assert e.value.errcode == "GDF_DATASET_EMPTY"


def test_mask_concat_with_device_resident_argument_arrays():
    sizes = [13, 64, 7]
    packed, bits = zip(*[G.rand_mask(n, 0.5) for n in sizes])
    dev = [torch.as_tensor(p).cuda() for p in packed]
    d_ptrs = torch.tensor([t.data_ptr() for t in dev], dtype=torch.int64, device="cuda")
    d_lens = torch.tensor(sizes, dtype=torch.int64, device="cuda")
    total = sum(sizes)
    out = torch.zeros((total + 7) // 8, dtype=torch.uint8, device="cuda")
    libgdf.gdf_mask_concat(ffi.cast("gdf_valid_type*", out.data_ptr()), total, ffi.cast("gdf_valid_type**", d_ptrs.data_ptr()),
                           ffi.cast("gdf_size_type*", d_lens.data_ptr()), len(sizes))
    torch.cuda.synchronize()
    got = np.unpackbits(out.cpu().numpy(), bitorder="little")[:total].astype(bool)
    np.testing.assert_array_equal(got, np.concatenate(bits))
