"""-m gpu: gdf_filter, gpu_comparison(_static) and gpu_apply_stencil through the C ABI vs the oracle.
Golden case from the reference (src/tests/cpp/sqls_tester.cu:82-172, baselines/sqls_tests_new_api.dat:69-71);
comparison matrix from src/tests/filterops_numeric/test_filterops.cu:79-177 (sizes 0-9, 6x6 dtypes)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import np_oracle
import oracle
from libgdf_b200 import columns as C
from libgdf_b200.libgdf_cffi import GDFError, ffi, libgdf
from gpu_utils import gen_rand, rand_mask

pytestmark = pytest.mark.gpu
GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.json")))
NP_TYPES = [np.int8, np.int16, np.int32, np.int64, np.float32, np.float64]
SUFFIX = {np.int8: "i8", np.int16: "i16", np.int32: "i32", np.int64: "i64", np.float32: "f32", np.float64: "f64"}


def gdf_filter(cols_np, vals):
    n = len(cols_np[0])
    cols = [C.column(c) for c in cols_np]
    structs = C.struct_array(cols)
    ncols = len(cols)
    d_cols = torch.zeros(ncols, dtype=torch.int64, device="cuda")
    d_types = torch.zeros(ncols, dtype=torch.int32, device="cuda")
    val_tensors = [torch.as_tensor(np.array([v], dtype=c.dtype)).cuda() for c, v in zip(cols_np, vals)]
    d_vals = torch.tensor([t.data_ptr() for t in val_tensors], dtype=torch.int64, device="cuda")
    d_indx = torch.full((max(n, 1),), -1, dtype=torch.int64, device="cuda")
    new_sz = ffi.new("size_t*")
    libgdf.gdf_filter(n, structs, ncols, ffi.cast("void**", d_cols.data_ptr()), ffi.cast("int*", d_types.data_ptr()),
                      ffi.cast("void**", d_vals.data_ptr()), ffi.cast("size_t*", d_indx.data_ptr()), new_sz)
    torch.cuda.synchronize()
    k = int(new_sz[0])
    # the caller's scratch arrays are part of the contract (reference sqls_ops.cu:27-41)
    assert d_cols.cpu().tolist() == [c.data.data_ptr() if len(c.data) else 0 for c in cols]
    assert d_types.cpu().tolist() == [int(c.cdata.dtype) for c in cols]
    return d_indx[:k].cpu().numpy().astype(np.uint64)


def test_filter_reference_golden():
    g = GOLDEN["filter"]
    cols = [np.array(c, dtype=t) for c, t in zip(g["cols"], g["dtypes"])]
    assert gdf_filter(cols, g["vals"]).tolist() == g["indices"]


@pytest.mark.parametrize("n", [0, 1, 31, 32, 33, 4095, 4096, 4097, 8193, 100_003, 1_048_576 + 17])
@pytest.mark.parametrize("np_t", NP_TYPES)
def test_filter_one_column(np_t, n):
    col = gen_rand(np_t, n, low=0, high=10) if np.dtype(np_t).kind != "f" else np.random.randint(0, 10, n).astype(np_t)
    got = gdf_filter([col], [3])
    want = np_oracle.filter_rows([col], [3])
    np.testing.assert_array_equal(got, want)          # bit-exact AND in ascending order


def test_filter_int64_matches_c_oracle_and_is_sorted():
    col = np.random.randint(0, 10, 3_000_000).astype(np.int64)
    got = gdf_filter([col], [3])
    np.testing.assert_array_equal(got, oracle.filter_i64(col, 3))
    assert (np.diff(got.astype(np.int64)) > 0).all()


def test_filter_extremes():
    n = 50_000
    np.testing.assert_array_equal(gdf_filter([np.full(n, 3, np.int64)], [3]), np.arange(n, dtype=np.uint64))
    assert len(gdf_filter([np.zeros(n, np.int64)], [3])) == 0
    nan_col = np.array([np.nan, 1.0, np.nan], dtype=np.float64)
    assert len(gdf_filter([nan_col], [np.nan])) == 0          # `!=` semantics: NaN rows never survive


def test_filter_many_columns_mixed_types():
    n = 200_003
    cols = [np.random.randint(0, 3, n).astype(np.int32), np.random.randint(0, 3, n).astype(np.int8),
            np.random.randint(0, 3, n).astype(np.float64), np.random.randint(0, 3, n).astype(np.int64)]
    vals = [1, 2, 0.0, 1]
    np.testing.assert_array_equal(gdf_filter(cols, vals), np_oracle.filter_rows(cols, vals))


def test_filter_unaligned_column():
    base = torch.as_tensor(np.random.randint(0, 10, 10_001).astype(np.int64)).cuda()
    view = base[1:]
    col = C.Column(view)
    host = view.cpu().numpy()
    structs = C.struct_array([col])
    d_cols = torch.zeros(1, dtype=torch.int64, device="cuda")
    d_types = torch.zeros(1, dtype=torch.int32, device="cuda")
    val = torch.tensor([3], dtype=torch.int64, device="cuda")
    d_vals = torch.tensor([val.data_ptr()], dtype=torch.int64, device="cuda")
    d_indx = torch.zeros(len(host), dtype=torch.int64, device="cuda")
    new_sz = ffi.new("size_t*")
    libgdf.gdf_filter(len(host), structs, 1, ffi.cast("void**", d_cols.data_ptr()), ffi.cast("int*", d_types.data_ptr()),
                      ffi.cast("void**", d_vals.data_ptr()), ffi.cast("size_t*", d_indx.data_ptr()), new_sz)
    np.testing.assert_array_equal(d_indx[: int(new_sz[0])].cpu().numpy(), np.nonzero(host == 3)[0])


def test_filter_rejects_mask():
    mask, _ = rand_mask(16)
    col = C.column(np.zeros(16, np.int32), mask)
    with pytest.raises(GDFError) as e:
        libgdf.gdf_filter(16, C.struct_array([col]), 1, ffi.NULL, ffi.NULL, ffi.NULL, ffi.NULL, ffi.new("size_t*"))
    assert e.value.errcode == "GDF_VALIDITY_UNSUPPORTED"


def _static_compare(lhs, value, value_t, op, lvalid=None):
    L = C.column(lhs, lvalid)
    n = len(lhs)
    O = C.empty_column(n, torch.int8, with_valid=True)
    getattr(libgdf, "gpu_comparison_static_" + SUFFIX[value_t])(L.cdata, value_t(value).item(), O.cdata, op)
    torch.cuda.synchronize()
    return O


@pytest.mark.parametrize("n", [0, 1, 5, 9, 1000, 65_537])
@pytest.mark.parametrize("value_t", NP_TYPES)
@pytest.mark.parametrize("lhs_t", NP_TYPES)
def test_comparison_static_equals_matrix(lhs_t, value_t, n):
    lhs = np.random.randint(0, 4, n).astype(lhs_t)
    O = _static_compare(lhs, 2, value_t, libgdf.GDF_EQUALS)
    np.testing.assert_array_equal(O.to_numpy(), np_oracle.comparison(lhs, value_t(2), libgdf.GDF_EQUALS))
    if n:
        assert (O.valid.cpu().numpy() == 0xFF).all() and O.cdata.null_count == 0


@pytest.mark.parametrize("op", ["GDF_EQUALS", "GDF_NOT_EQUALS", "GDF_LESS_THAN", "GDF_LESS_THAN_OR_EQUALS",
                                "GDF_GREATER_THAN", "GDF_GREATER_THAN_OR_EQUALS"])
def test_comparison_static_all_operators_i64(op):
    lhs = np.random.randint(-5, 5, 100_001).astype(np.int64)
    O = _static_compare(lhs, 1, np.int64, getattr(libgdf, op))
    np.testing.assert_array_equal(O.to_numpy(), np_oracle.comparison(lhs, np.int64(1), getattr(libgdf, op)))


def test_comparison_static_propagates_mask():
    n = 1003
    lhs = np.random.randint(0, 4, n).astype(np.int32)
    mask, bits = rand_mask(n)
    O = _static_compare(lhs, 2, np.int32, libgdf.GDF_EQUALS, lvalid=mask)
    np.testing.assert_array_equal(O.valid.cpu().numpy(), mask)
    assert O.cdata.null_count == n - bits.sum()


@pytest.mark.parametrize("rhs_t", NP_TYPES)
@pytest.mark.parametrize("lhs_t", NP_TYPES)
def test_comparison_columns_matrix(lhs_t, rhs_t):
    n = 1237
    lhs, rhs = np.random.randint(0, 4, n).astype(lhs_t), np.random.randint(0, 4, n).astype(rhs_t)
    lm, lb = rand_mask(n)
    rm, rb = rand_mask(n)
    L, R = C.column(lhs, lm), C.column(rhs, rm)
    O = C.empty_column(n, torch.int8, with_valid=True)
    libgdf.gpu_comparison(L.cdata, R.cdata, O.cdata, libgdf.GDF_GREATER_THAN)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(O.to_numpy(), np_oracle.comparison(lhs, rhs, libgdf.GDF_GREATER_THAN))
    np.testing.assert_array_equal(O.valid.cpu().numpy(), lm & rm)
    assert O.cdata.null_count == n - (lb & rb).sum()


@pytest.mark.parametrize("n", [1, 7, 8, 9, 4096, 8191, 8192, 8193, 300_007])
@pytest.mark.parametrize("np_t", [np.int8, np.int16, np.int32, np.int64, np.float32, np.float64])
def test_apply_stencil(np_t, n):
    data = gen_rand(np_t, n)
    stencil = (np.random.rand(n) < 0.1).astype(np.int8)
    svalid = np.full((n + 7) // 8, 0xFF, np.uint8)
    L = C.column(data)
    S = C.column(stencil, svalid)
    O = C.empty_column(n, L.data.dtype, with_valid=True)
    libgdf.gpu_apply_stencil(L.cdata, S.cdata, O.cdata)
    torch.cuda.synchronize()
    want, want_mask = np_oracle.apply_stencil(data, stencil, svalid)
    assert O.size == len(want)
    np.testing.assert_array_equal(O.to_numpy(), want)
    np.testing.assert_array_equal(O.valid.cpu().numpy(), want_mask)


def test_apply_stencil_reads_mask_msb_first():
    n = 20_000
    data = np.arange(n, dtype=np.int64)
    stencil = np.ones(n, np.int8)
    svalid, _ = rand_mask(n)
    L, S = C.column(data), C.column(stencil, svalid)
    O = C.empty_column(n, torch.int64, with_valid=True)
    libgdf.gpu_apply_stencil(L.cdata, S.cdata, O.cdata)
    torch.cuda.synchronize()
    want, _ = np_oracle.apply_stencil(data, stencil, svalid)
    np.testing.assert_array_equal(O.to_numpy(), want)


def test_filter_then_stencil_chain_c2_shape():
    """BASELINE config C2 at test size: int64 uniform [0,10), == 3, both API routes agree."""
    n = 2_000_000
    col = np.random.randint(0, 10, n).astype(np.int64)
    idx = gdf_filter([col], [3])
    O = _static_compare(col, 3, np.int64, libgdf.GDF_EQUALS)
    L = C.column(col)
    out = C.empty_column(n, torch.int64, with_valid=True)
    libgdf.gpu_apply_stencil(L.cdata, O.cdata, out.cdata)
    torch.cuda.synchronize()
    assert out.size == len(idx)
    np.testing.assert_array_equal(out.to_numpy(), col[idx.astype(np.int64)])
    assert abs(len(idx) / n - 0.1) < 0.01


def test_stencil_errors():
    L = C.column(np.zeros(8, np.int32), np.full(1, 0xFF, np.uint8))
    S = C.column(np.zeros(8, np.int8), np.full(1, 0xFF, np.uint8))
    O = C.empty_column(8, torch.int32, with_valid=True)
    with pytest.raises(GDFError) as e:
        libgdf.gpu_apply_stencil(L.cdata, S.cdata, O.cdata)
    assert e.value.errcode == "GDF_VALIDITY_UNSUPPORTED"


# ---- forward progress of the streaming select ----
def _filter_big(n, seed=3):
    col = torch.randint(0, 10, (n,), generator=torch.Generator(device="cuda").manual_seed(seed), device="cuda", dtype=torch.int64)
    c = C.Column(col)
    structs = C.struct_array([c])
    d_cols = torch.zeros(1, dtype=torch.int64, device="cuda")
    d_types = torch.zeros(1, dtype=torch.int32, device="cuda")
    val = torch.tensor([3], dtype=torch.int64, device="cuda")
    d_vals = torch.tensor([val.data_ptr()], dtype=torch.int64, device="cuda")
    d_indx = torch.empty(n, dtype=torch.int64, device="cuda")
    new_sz = ffi.new("size_t*")

    def run():
        libgdf.gdf_filter(n, structs, 1, ffi.cast("void**", d_cols.data_ptr()), ffi.cast("int*", d_types.data_ptr()),
                          ffi.cast("void**", d_vals.data_ptr()), ffi.cast("size_t*", d_indx.data_ptr()), new_sz)
        return d_indx[: int(new_sz[0])].clone()
    want = torch.nonzero(col == 3).flatten()
    return run, want


@pytest.mark.parametrize("mode", [0, 1])
def test_filter_both_chunk_dealings_agree_at_many_chunks_per_cta(mode):
    """40 M rows = 611 chunks for at most 296 persistent CTAs: every CTA takes several chunks, so the static dealing
    (cooperative launch) and the ticket dealing both run their steady state.  Bit-exact and ascending."""
    run, want = _filter_big(40_000_003)
    prev = libgdf.gdfx_set_select_dealing(mode)
    try:
        got = run()
    finally:
        libgdf.gdfx_set_select_dealing(prev)
    assert torch.equal(got, want)


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("held_sms", [74, 140])
def test_filter_completes_while_a_foreign_kernel_holds_sms(mode, held_sms):
    """A foreign kernel parks CTAs that fill `held_sms` SMs for ~40 ms on another (non-blocking) stream, so the select's
    persistent grid cannot be co-resident while it runs.  Mode 0: the cooperative launch is simply not started until
    the grid fits; mode 1: the ticket dealing makes progress on whatever SMs are free.  Neither may hang or mis-rank."""
    run, want = _filter_big(30_000_001, seed=9)
    run()                                            # warm: scratch allocation, attribute calls
    prev = libgdf.gdfx_set_select_dealing(mode)
    try:
        libgdf.gdfx_debug_occupy_sms(held_sms, 40_000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        got = run()
        e1.record()
        torch.cuda.synchronize()
    finally:
        libgdf.gdfx_set_select_dealing(prev)
    assert torch.equal(got, want)
    print("mode %d, %d SMs held: gdf_filter returned after %.2f ms" % (mode, held_sms, e0.elapsed_time(e1)))
