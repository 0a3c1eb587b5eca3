"""-m gpu: hash group-by through the C ABI vs the C oracle.
Case list follows the reference's gtest (src/tests/groupby/groupby-test.cu:369-445: one key x 8,
AllKeysSame 2^14, AllKeysDifferent 2^14, WarpKeysSame 2^10x32, BlockKeysSame 2^10x256, EmptyInput) and
its parameter matrix (test_parameters.cuh:126-153: MIN/MAX/SUM/COUNT/AVG, 1-3 key columns, value types
int32/int64/float/double ...).  Comparison is order-independent (the reference's is, :346-364):
exact for integral outputs, 1e-6 / 1e-4 relative for float64 / float32 (the reference allows 1 %)."""
import json
import os

import numpy as np
import pytest
import torch

import oracle
from libgdf_b200 import columns as C
from libgdf_b200.libgdf_cffi import GDFError, ffi, libgdf
from gpu_utils import gen_rand, groupby, rows_as_sorted_tuples, rand_mask

pytestmark = pytest.mark.gpu
GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.json")))
OPS = {"sum": oracle.OP_SUM, "min": oracle.OP_MIN, "max": oracle.OP_MAX, "avg": oracle.OP_AVG, "count": oracle.OP_COUNT}


def check(op, keys, values, out_np_dtype=None):
    out_np_dtype = np.dtype(out_np_dtype or values.dtype)
    gk, ga = groupby(op, keys, values, out_np_dtype)
    ok, oa = oracle.groupby(OPS[op], keys, values, out_dtype=oracle.NP_TO_GDF[out_np_dtype])
    assert len(ga) == len(oa), "group count"
    got, want = rows_as_sorted_tuples(gk, ga), rows_as_sorted_tuples(ok, oa)
    assert len(set(r[:-1] for r in got)) == len(got), "duplicate groups"
    if ga.dtype.kind == "f":
        assert [r[:-1] for r in got] == [r[:-1] for r in want]
        rtol = 1e-4 if ga.dtype == np.float32 else 1e-6
        np.testing.assert_allclose([r[-1] for r in got], [r[-1] for r in want], rtol=rtol, atol=rtol)
    else:
        assert got == want


@pytest.mark.parametrize("op", ["sum", "min", "max", "count", "avg"])
def test_reference_golden_groups(op):
    g = GOLDEN["groupby"]
    keys = [np.array(c, dtype=t) for c, t in zip(g["keys"], g["key_dtypes"])]
    vals = np.array(g["agg_min_max" if op in ("min", "max") else "agg_sum_count_avg"], dtype=np.float64)
    out_t = np.int32 if op == "count" else np.float64
    gk, ga = groupby(op, keys, vals, out_t)
    got = rows_as_sorted_tuples(gk, ga)
    want = sorted(zip(*g["group_keys_sorted"], g[op]))
    assert got == want


@pytest.mark.parametrize("val_t", [np.int32, np.int64, np.float32, np.float64])
@pytest.mark.parametrize("op", ["sum", "min", "max"])
@pytest.mark.parametrize("shape", ["eight", "all_same", "all_different", "warp_same", "block_same"])
def test_reference_case_matrix_one_key(shape, op, val_t):
    if shape == "eight":
        keys = np.array([0, 1, 2, 3, 0, 1, 2, 3], np.int64)
    elif shape == "all_same":
        keys = np.zeros(1 << 14, np.int64)
    elif shape == "all_different":
        keys = np.random.permutation(1 << 14).astype(np.int64)
    elif shape == "warp_same":
        keys = np.repeat(np.arange(1 << 10), 32).astype(np.int64)
    else:
        keys = np.repeat(np.arange(1 << 10), 256).astype(np.int64)
    vals = gen_rand(val_t, len(keys), 0, 1000)
    check(op, [keys], vals)


@pytest.mark.parametrize("key_t", [np.int8, np.int16, np.int32, np.int64, np.float32, np.float64])
def test_key_types(key_t):
    n = 20_000
    keys = np.random.randint(0, 100, n).astype(key_t)
    check("sum", [keys], gen_rand(np.int64, n, 0, 1000))
    check("count", [keys], gen_rand(np.int64, n, 0, 1000), np.int32)


@pytest.mark.parametrize("op", ["sum", "min", "max", "count", "avg"])
def test_multi_column_keys(op):
    n = 100_003
    keys = [np.random.randint(0, 30, n).astype(np.int32), np.random.randint(0, 7, n).astype(np.int64),
            np.random.randint(0, 3, n).astype(np.float64)]
    vals = gen_rand(np.float64, n) if op != "count" else gen_rand(np.int32, n)
    check(op, keys, vals, np.int64 if op == "count" else np.float64)


@pytest.mark.parametrize("val_t,out_t", [(np.int8, np.int8), (np.int16, np.int16), (np.int32, np.int32)])
def test_narrow_sums_wrap_like_the_reference(val_t, out_t):
    n = 70_000
    keys = np.random.randint(0, 5, n).astype(np.int32)
    vals = gen_rand(val_t, n, 0, 120)
    check("sum", [keys], vals, out_t)


@pytest.mark.parametrize("val_t,out_t", [(np.int32, np.float64), (np.int64, np.int64), (np.float32, np.float32),
                                         (np.float64, np.float64), (np.int64, np.float32), (np.int32, np.int32)])
def test_avg_type_matrix(val_t, out_t):
    n = 50_000
    keys = np.random.randint(0, 200, n).astype(np.int64)
    vals = gen_rand(val_t, n, 0, 1000)
    check("avg", [keys], vals, out_t)


@pytest.mark.parametrize("out_t", [np.int32, np.int64, np.float32, np.float64])
def test_count_output_types(out_t):
    n = 30_000
    keys = np.random.randint(0, 50, n).astype(np.int32)
    check("count", [keys], gen_rand(np.float64, n), out_t)


def test_c4_shape_zipf_int64():
    """BASELINE config C4 at test size: Zipf(1.05) ranks over 1e5 ids mapped through a permutation,
    int64 values in [0,1000), sum; exact multiset parity + conservation of the grand total."""
    n, groups = 4_000_000, 100_000
    cdf = np.cumsum(1.0 / np.arange(1, groups + 1) ** 1.05)
    cdf /= cdf[-1]
    ranks = np.searchsorted(cdf, np.random.rand(n))
    ids = (np.random.permutation(groups).astype(np.int64) * 7919 + 13)
    keys = ids[ranks]
    vals = np.random.randint(0, 1000, n).astype(np.int64)
    check("sum", [keys], vals)
    gk, ga = groupby("sum", [keys], vals)
    assert int(ga.sum()) == int(vals.sum()) and len(np.unique(gk[0])) == len(gk[0])


def test_more_groups_than_level1_table_spills_to_level2():
    """> 2^22 distinct keys forces the bounded level-1 table to spill (groupby.cu two-level scheme)."""
    n = 6_000_000
    keys = (np.random.permutation(n).astype(np.int64) * 3 + 1)
    keys[:1000] = keys[1000:2000]                    # a few duplicates across the spill boundary
    vals = np.random.randint(0, 1000, n).astype(np.int64)
    gk, ga = groupby("sum", [keys], vals)
    assert len(gk[0]) == len(np.unique(keys)) and int(ga.sum()) == int(vals.sum())
    order = np.argsort(gk[0])
    uk, inv = np.unique(keys, return_inverse=True)
    want = np.bincount(inv, weights=None, minlength=len(uk))  # counts
    sums = np.zeros(len(uk), np.int64)
    np.add.at(sums, inv, vals)
    np.testing.assert_array_equal(gk[0][order], uk)
    np.testing.assert_array_equal(ga[order], sums)


def test_generic_path_spill_multi_column():
    n = 5_000_000
    k0 = np.random.permutation(n).astype(np.int32)
    k1 = np.zeros(n, np.int8)
    vals = np.ones(n, np.int32)
    gk, ga = groupby("count", [k0, k1], vals, np.int32)
    assert len(ga) == n and (ga == 1).all()
    np.testing.assert_array_equal(np.sort(gk[0]), np.arange(n, dtype=np.int32))


def test_key_equal_to_empty_marker():
    keys = np.array([-1, 5, -1, -1, 7, 5], np.int64)      # -1 == all-ones bit pattern
    vals = np.array([1, 2, 3, 4, 5, 6], np.int64)
    check("sum", [keys], vals)
    check("avg", [keys], vals, np.float64)


def test_empty_input_and_errors():
    ctx = ffi.new("gdf_context*")
    libgdf.gdf_context_view(ctx, 0, libgdf.GDF_HASH, 0, 0, 0)
    k, v = C.column(np.zeros(0, np.int32)), C.column(np.zeros(0, np.int32))
    ok, oa = C.empty_column(4, torch.int32), C.empty_column(4, torch.int32)
    libgdf.gdf_group_by_sum(1, C.column_array([k]), v.cdata, ffi.NULL, C.column_array([ok]), oa.cdata, ctx)
    assert oa.size == 0 and ok.size == 0
    mask, _ = rand_mask(8)
    km = C.column(np.zeros(8, np.int32), mask)
    v8 = C.column(np.zeros(8, np.int32))
    with pytest.raises(GDFError) as e:
        libgdf.gdf_group_by_sum(1, C.column_array([km]), v8.cdata, ffi.NULL, C.column_array([ok]), oa.cdata, ctx)
    assert e.value.errcode == "GDF_VALIDITY_UNSUPPORTED"
    with pytest.raises(GDFError) as e:
        libgdf.gdf_group_by_sum(1, C.column_array([v8]), v8.cdata, ffi.NULL, C.column_array([ok]), oa.cdata, ffi.NULL)
    assert e.value.errcode == "GDF_DATASET_EMPTY"
