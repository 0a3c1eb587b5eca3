"""-m gpu: three-way parity  reference kernels (oracle/_ref/libgdf_ref.so)  ==  CPU oracle  ==  product.

oracle/_ref/libgdf_ref.so is gpuopenanalytics/libgdf rebuilt for sm_100a by oracle/build_ref.sh (test
infrastructure; it travels to the GPU box as a prebuilt file - /root/reference is never read here).
This file is what PINS the CPU oracle: every oracle function the other GPU tests rely on is checked
against the reference's own kernels on the same seeded inputs, and the product library is checked
against both.  Comparison rules are the reference's own (join-tests.cu:341-345 sort both sides;
groupby-test.cu:346-364 order-independent lookup): bit-exact for integers / indices, multiset equality
where the reference leaves the order unspecified, 1e-9 relative for float64 sums.

Operators the reference implements with documented bugs are compared on the non-buggy subset only
(SURVEY.md section 8a "quirk decisions"): gpu_comparison LESS_THAN/LESS_THAN_OR_EQUALS are skipped."""
import os

import cffi
import numpy as np
import pytest
import torch

import oracle
from oracle import np_oracle
from libgdf_b200 import columns as C
from libgdf_b200._cdef import header_cdef
from libgdf_b200.libgdf_cffi import ffi, libgdf
import gpu_utils as G

pytestmark = pytest.mark.gpu
REF_SO = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libgdf_ref.so")


@pytest.fixture(scope="module")
def ref():
    """The reference library, dlopen'ed RTLD_LOCAL next to ours (it was linked -Bsymbolic, so its
    gdf_*/rmm* calls bind inside itself).  Structs are passed between the two ffi objects by address."""
    if not os.path.isfile(REF_SO):
        pytest.skip("oracle/_ref/libgdf_ref.so not built (run oracle/build_ref.sh where /root/reference exists)")
    rffi = cffi.FFI()
    rffi.cdef(header_cdef("gdf/cffi/types.h", "gdf/cffi/functions.h", "memory.h"))
    lib = rffi.dlopen(REF_SO)
    opts = rffi.new("rmmOptions_t*")
    opts.allocation_mode = lib.CudaDefaultAllocation
    opts.initial_pool_size = 0
    opts.enable_logging = False
    assert lib.rmmInitialize(opts) == 0
    return _RefApi(rffi, lib)


class _RefApi(object):
    """Gives the reference library the same call surface gpu_utils expects from `libgdf`: arguments
    built with OUR ffi (gdf_column*, gdf_column*[], int[], gdf_context*) are re-cast by address."""

    def __init__(self, rffi, lib):
        self.rffi, self.lib = rffi, lib

    def __getattr__(self, name):
        attr = getattr(self.lib, name)
        if not callable(attr):
            return attr
        rffi = self.rffi
        argtypes = rffi.typeof(attr).args

        def call(*args):
            conv = []
            for a, t in zip(args, argtypes):
                if isinstance(a, ffi.CData) and ffi.typeof(a).kind in ("pointer", "array"):
                    conv.append(rffi.cast(t, int(ffi.cast("uintptr_t", a))))
                else:
                    conv.append(a)
            return attr(*conv)
        return call


def ref_mask(n, p=0.5):
    bits = np.ones(n, dtype=bool)
    bits[n // 2:] = np.random.rand(n - n // 2) < p
    return np.packbits(bits, bitorder="little")


# ------------------------------------------------------------------------------------------------
# row hash: pins the MurmurHash3 / hash_combine restatement VALUE BY VALUE (the reference's own tests
# only check "equal rows hash equal", hash-test.cu:157)
# ------------------------------------------------------------------------------------------------
def _hash(api, cols_np, func="GDF_HASH_MURMUR3"):
    cols = [C.column(c, api=api) for c in cols_np]
    out = C.empty_column(len(cols_np[0]), torch.int32, api=api)
    rc = api.gdf_hash(len(cols), C.column_array(cols), getattr(api, func), out.cdata)
    assert rc in (None, 0)
    torch.cuda.synchronize()
    return out.to_numpy()


@pytest.mark.parametrize("np_t", [np.int8, np.int16, np.int32, np.int64, np.float32, np.float64])
def test_hash_values_single_column(ref, np_t):
    col = G.gen_rand(np_t, 20_011)
    want = _hash(ref, [col])
    np.testing.assert_array_equal(oracle.hash_rows([col]), want)
    np.testing.assert_array_equal(_hash(libgdf, [col]), want)


def test_hash_values_known_answers_and_multi_column(ref):
    k0 = np.array([28467476447, 0, 1, -1, 42], dtype=np.int64)
    k1 = np.array([28, 0, 1, 28, 28], dtype=np.int32)
    want = _hash(ref, [k0, k1])
    assert (int(want[0]) & 0xffffffff) == 0x29ec6a17          # SURVEY.md 8c offline KAT, now pinned by the reference
    np.testing.assert_array_equal(oracle.hash_rows([k0, k1]), want)
    np.testing.assert_array_equal(_hash(libgdf, [k0, k1]), want)
    single = _hash(ref, [k0])
    assert [int(x) & 0xffffffff for x in single[1:]] == [0x63852afc, 0x53075d44, 0x627564e8, 0x6f8f913e]
    n = 30_000
    cols = [G.gen_rand(np.int64, n, 0, 50), G.gen_rand(np.int32, n, 0, 4), G.gen_rand(np.float64, n), G.gen_rand(np.int8, n)]
    want = _hash(ref, cols)
    np.testing.assert_array_equal(oracle.hash_rows(cols), want)
    np.testing.assert_array_equal(_hash(libgdf, cols), want)


# ------------------------------------------------------------------------------------------------
# hash partition: same partition sizes, same multiset of rows inside every partition
# ------------------------------------------------------------------------------------------------
def _partition(api, cols_np, hash_idx, nparts):
    n = len(cols_np[0])
    cols = [C.column(c, api=api) for c in cols_np]
    outs = [C.empty_column(n, c.data.dtype, api=api) for c in cols]
    offsets = ffi.new("int[]", nparts)
    rc = api.gdf_hash_partition(len(cols), C.column_array(cols), ffi.new("int[]", hash_idx), len(hash_idx), nparts,
                                C.column_array(outs), offsets, api.GDF_HASH_MURMUR3)
    assert rc in (None, 0)
    torch.cuda.synchronize()
    return [o.to_numpy() for o in outs], list(offsets)


@pytest.mark.parametrize("nparts", [1, 5, 8, 257])
def test_hash_partition(ref, nparts):
    n = 100_003
    cols = [G.gen_rand(np.int64, n, 0, 1 << 30), G.gen_rand(np.int32, n)]
    r_out, r_off = _partition(ref, cols, [0], nparts)
    g_out, g_off = _partition(libgdf, cols, [0], nparts)
    assert g_off == r_off
    np.testing.assert_array_equal(np.bincount(oracle.partition_ids([cols[0]], nparts), minlength=nparts),
                                  np.diff(r_off + [n]))
    bounds = r_off + [n]
    for p in range(nparts):
        lo, hi = bounds[p], bounds[p + 1]
        assert sorted(zip(r_out[0][lo:hi].tolist(), r_out[1][lo:hi].tolist())) == \
            sorted(zip(g_out[0][lo:hi].tolist(), g_out[1][lo:hi].tolist()))


# ------------------------------------------------------------------------------------------------
# joins
# ------------------------------------------------------------------------------------------------
JOIN_KEYSETS = [[np.int64], [np.int32], [np.float64], [np.int64, np.int32], [np.int32, np.float64, np.int64]]


@pytest.mark.parametrize("kind", ["inner", "left", "full"])
@pytest.mark.parametrize("key_types", JOIN_KEYSETS, ids=lambda ts: "-".join(np.dtype(t).name for t in ts))
@pytest.mark.parametrize("masks", [False, True], ids=["nomask", "mask"])
def test_join(ref, key_types, kind, masks):
    nl, nr = 20_000, 6_000
    rng = 3000 if len(key_types) == 1 else 14
    l = [np.random.randint(0, rng, nl).astype(t) for t in key_types]
    r = [np.random.randint(0, rng, nr).astype(t) for t in key_types]
    lv = [ref_mask(nl) for _ in key_types] if masks else None
    rv = [ref_mask(nr) for _ in key_types] if masks else None
    rl, rr = G.join(kind, l, r, lv, rv, api=ref)
    want = G.sorted_pairs(rl, rr)
    ol, orr = oracle.join({"inner": oracle.JOIN_INNER, "left": oracle.JOIN_LEFT, "full": oracle.JOIN_FULL}[kind], l, r, lv, rv)
    np.testing.assert_array_equal(G.sorted_pairs(ol, orr), want)
    gl, gr = G.join(kind, l, r, lv, rv)
    np.testing.assert_array_equal(G.sorted_pairs(gl, gr), want)


@pytest.mark.parametrize("kind", ["inner", "left"])
def test_join_partitioned_path_int64(ref, kind):
    """Large enough for the product's radix-partitioned path (build > 2^20 rows), C3-shaped: unique
    build keys, probe keys partly missing, 30 % NULL probe rows for LEFT (C5's null rule)."""
    nb, npr = 1_300_000, 2_600_000
    build = np.random.permutation(nb).astype(np.int64)
    probe = np.random.randint(0, 2 * nb, npr).astype(np.int64)
    pv = [np.packbits(np.random.rand(npr) < 0.7, bitorder="little")] if kind == "left" else None
    rl, rr = G.join(kind, [probe], [build], pv, None, api=ref)
    gl, gr = G.join(kind, [probe], [build], pv, None)
    np.testing.assert_array_equal(G.sorted_pairs(gl, gr), G.sorted_pairs(rl, rr))


# ------------------------------------------------------------------------------------------------
# group-by (HASH)
# ------------------------------------------------------------------------------------------------
OPS = {"sum": oracle.OP_SUM, "min": oracle.OP_MIN, "max": oracle.OP_MAX, "count": oracle.OP_COUNT, "avg": oracle.OP_AVG}


def _close_rows(a, b, rel):
    assert len(a) == len(b)
    for x, y in zip(a, b):
        assert x[:-1] == y[:-1]
        assert abs(x[-1] - y[-1]) <= rel * max(1.0, abs(y[-1]))


@pytest.mark.parametrize("op", ["sum", "min", "max", "count", "avg"])
@pytest.mark.parametrize("key_types,val_t", [([np.int64], np.int64), ([np.int32], np.int32), ([np.int64, np.int32], np.int64),
                                             ([np.int32, np.int8, np.int64], np.float64), ([np.int64], np.float64)],
                         ids=["i64-i64", "i32-i32", "i64i32-i64", "i32i8i64-f64", "i64-f64"])
def test_groupby(ref, key_types, val_t, op):
    n = 40_000
    keys = [np.random.randint(0, 300 if len(key_types) == 1 else 7, n).astype(t) for t in key_types]
    vals = G.gen_rand(val_t, n, 0, 1000)
    out_t = val_t
    rk, ra = G.groupby(op, keys, vals, out_t, api=ref)
    ok, oa = oracle.groupby(OPS[op], keys, vals)
    gk, ga = G.groupby(op, keys, vals, out_t)
    want = G.rows_as_sorted_tuples(rk, ra)
    if np.dtype(val_t).kind == "f" and op in ("sum", "avg"):
        # float accumulation order differs between atomics and the CPU loop: 1e-9 relative (float64)
        _close_rows(G.rows_as_sorted_tuples(ok, oa), want, 1e-9)
        _close_rows(G.rows_as_sorted_tuples(gk, ga), want, 1e-9)
    else:
        assert G.rows_as_sorted_tuples(ok, oa) == want
        assert G.rows_as_sorted_tuples(gk, ga) == want


def test_groupby_zipf_sum_int64(ref):
    """C4-shaped (Zipf s=1.05 keys, int64 values in [0,1000)) at a size the reference handles (< 2^29 rows)."""
    n, groups = 2_000_000, 20_000
    ranks = np.arange(1, groups + 1, dtype=np.float64) ** -1.05
    cdf = np.cumsum(ranks) / ranks.sum()
    ids = np.random.permutation(groups).astype(np.int64) * 7919 + 13
    keys = ids[np.minimum(np.searchsorted(cdf, np.random.rand(n)), groups - 1)]
    vals = np.random.randint(0, 1000, n).astype(np.int64)
    rk, ra = G.groupby("sum", [keys], vals, api=ref)
    gk, ga = G.groupby("sum", [keys], vals)
    ok, oa = oracle.groupby(oracle.OP_SUM, [keys], vals)
    want = G.rows_as_sorted_tuples(rk, ra)
    assert G.rows_as_sorted_tuples(gk, ga) == want
    assert G.rows_as_sorted_tuples(ok, oa) == want


# ------------------------------------------------------------------------------------------------
# gdf_filter, comparison + apply_stencil, reductions, add
# ------------------------------------------------------------------------------------------------
def _filter(api, cols_np, vals):
    n = len(cols_np[0])
    cols = [C.column(c, api=api) for c in cols_np]
    d_cols = torch.zeros(len(cols), dtype=torch.int64, device="cuda")
    d_types = torch.zeros(len(cols), dtype=torch.int32, device="cuda")
    val_t = [torch.as_tensor(np.array([v], dtype=c.dtype)).cuda() for c, v in zip(cols_np, vals)]
    d_vals = torch.tensor([t.data_ptr() for t in val_t], dtype=torch.int64, device="cuda")
    d_indx = torch.full((max(n, 1),), -1, dtype=torch.int64, device="cuda")
    new_sz = ffi.new("size_t*")
    rc = api.gdf_filter(n, C.struct_array(cols), len(cols), ffi.cast("void**", d_cols.data_ptr()),
                        ffi.cast("int*", d_types.data_ptr()), ffi.cast("void**", d_vals.data_ptr()),
                        ffi.cast("size_t*", d_indx.data_ptr()), new_sz)
    assert rc in (None, 0)
    torch.cuda.synchronize()
    return d_indx[: int(new_sz[0])].cpu().numpy().astype(np.uint64)


def test_filter(ref):
    n = 300_007
    col = np.random.randint(0, 10, n).astype(np.int64)
    want = _filter(ref, [col], [3])
    np.testing.assert_array_equal(np_oracle.filter_rows([col], [3]), want)
    np.testing.assert_array_equal(oracle.filter_i64(col, 3), want)
    np.testing.assert_array_equal(_filter(libgdf, [col], [3]), want)
    cols = [np.random.randint(0, 3, n).astype(np.int32), np.random.randint(0, 3, n).astype(np.float64)]
    want = _filter(ref, cols, [1, 2.0])
    np.testing.assert_array_equal(np_oracle.filter_rows(cols, [1, 2.0]), want)
    np.testing.assert_array_equal(_filter(libgdf, cols, [1, 2.0]), want)


def _compare_and_compact(api, data, value, op):
    n = len(data)
    L = C.column(data, api=api)
    S = C.empty_column(n, torch.int8, with_valid=True, api=api)
    rc = api.gpu_comparison_static_i64(L.cdata, int(value), S.cdata, op)
    assert rc in (None, 0)
    O = C.empty_column(n, torch.int64, with_valid=True, api=api)
    rc = api.gpu_apply_stencil(L.cdata, S.cdata, O.cdata)
    assert rc in (None, 0)
    torch.cuda.synchronize()
    return S.data.cpu().numpy(), O.data[: int(O.cdata.size)].cpu().numpy()


@pytest.mark.parametrize("op_name", ["GDF_EQUALS", "GDF_NOT_EQUALS", "GDF_GREATER_THAN", "GDF_GREATER_THAN_OR_EQUALS"])
def test_comparison_static_then_apply_stencil(ref, op_name):
    n = 100_000          # multiple of 8: the reference packs a ragged last mask byte MSB-first (documented quirk)
    data = np.random.randint(0, 10, n).astype(np.int64)
    rs, ro = _compare_and_compact(ref, data, 3, getattr(ref, op_name))
    gs, go = _compare_and_compact(libgdf, data, 3, getattr(libgdf, op_name))
    code = {"GDF_EQUALS": 0, "GDF_NOT_EQUALS": 1, "GDF_GREATER_THAN": 4, "GDF_GREATER_THAN_OR_EQUALS": 5}[op_name]
    np.testing.assert_array_equal(np_oracle.comparison(data, np.int64(3), code), rs)
    np.testing.assert_array_equal(gs, rs)
    np.testing.assert_array_equal(go, ro)
    kept, _ = np_oracle.apply_stencil(data, rs, np.full((n + 7) // 8, 0xFF, np.uint8))
    np.testing.assert_array_equal(kept, ro)


def _reduce(api, name, data, valid=None):
    col = C.column(data, valid, api=api)
    out = torch.zeros(128, dtype=col.data.dtype, device="cuda")
    suffix, ctype = {np.dtype(np.int64): ("i64", "int64_t*"), np.dtype(np.int32): ("i32", "int32_t*"),
                     np.dtype(np.float64): ("f64", "double*")}[data.dtype]
    rc = getattr(api, "gdf_%s_%s" % (name, suffix))(col.cdata, ffi.cast(ctype, out.data_ptr()), 128)
    assert rc in (None, 0)
    torch.cuda.synchronize()
    return out[0].cpu().numpy()


@pytest.mark.parametrize("name", ["sum", "min", "max"])
@pytest.mark.parametrize("np_t", [np.int64, np.int32, np.float64])
@pytest.mark.parametrize("masked", [False, True])
def test_reductions(ref, name, np_t, masked):
    # 16000 rows: every one of the reference's 128 blocks runs its block-reduce loop exactly once.  With
    # more rows the reference re-enters cub::BlockReduce on the same temp_storage without a barrier
    # (reductions.cu:40-55), a race that produced a wrong int32 sum on sm_100a at 200003 rows in this
    # suite's first GPU run; larger sizes are covered against the oracle in test_reductions_gpu.py.
    n = 16_000
    data = G.gen_rand(np_t, n)
    valid = G.rand_mask(n)[0] if masked else None
    want = _reduce(ref, name, data, valid)
    got, orc = _reduce(libgdf, name, data, valid), np_oracle.reduce(name, data, valid)
    if np.dtype(np_t).kind == "f" and name == "sum":
        np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(orc, want, rtol=1e-9, atol=1e-9)
    else:
        assert got == want and orc == want


def test_add_config_c1(ref):
    """BASELINE config 1: gdf_add on two 1M-row int32 columns."""
    a, b = G.gen_rand(np.int32, 1_000_000), G.gen_rand(np.int32, 1_000_000)
    res = []
    for api in (ref, libgdf):
        A, B, O = C.column(a, api=api), C.column(b, api=api), C.column(np.zeros_like(a), api=api)
        rc = api.gdf_add_generic(A.cdata, B.cdata, O.cdata)
        assert rc in (None, 0)
        torch.cuda.synchronize()
        res.append(O.to_numpy())
    np.testing.assert_array_equal(res[1], res[0])
    np.testing.assert_array_equal(np_oracle.binary_op("add", a, b, np.zeros_like(a)), res[0])


# ------------------------------------------------------------------------------------------------
# result_cols materialisation: LEFT / FULL rows with index -1, masked payload columns
# ------------------------------------------------------------------------------------------------
def _join_result_cols(api, kind, lk, lp, lpv, rk, rp, rpv):
    """Tables: left = [payload int32 (masked), key int64], right = [key int64, payload float64 (masked)], join on the key.
    Returns one tuple per output row, undefined cells (validity bit 0) as None:
    (left payload, key, right payload, left index, right index)."""
    L = [C.column(lp, lpv, api=api), C.column(lk, api=api)]
    R = [C.column(rk, api=api), C.column(rp, rpv, api=api)]
    res = [ffi.new("gdf_column*") for _ in range(3)]
    res_arr = ffi.new("gdf_column*[]", res)
    ctx = ffi.new("gdf_context*")
    api.gdf_context_view(ctx, 0, api.GDF_HASH, 0, 0, 0)
    out_l, out_r = ffi.new("gdf_column*"), ffi.new("gdf_column*")
    rc = getattr(api, "gdf_%s_join" % kind)(C.column_array(L), 2, ffi.new("int[]", [1]), C.column_array(R), 2, ffi.new("int[]", [0]),
                                             1, 3, res_arr, out_l, out_r, ctx)
    assert rc in (None, 0)
    torch.cuda.synchronize()
    n = int(out_l.size)
    assert all(int(c.size) == n for c in res)

    def host(col, np_t):
        data = C.alias_column_data(col, np_t).cpu().numpy()
        if col.valid == ffi.NULL:
            return data, np.ones(n, bool)
        vb = C._alias(int(ffi.cast("uintptr_t", col.valid)), (n + 7) // 8, np.uint8).cpu().numpy()
        return data, np.unpackbits(vb, bitorder="little")[:n].astype(bool)
    li, ri = C.alias_column_data(out_l, np.int32).cpu().numpy(), C.alias_column_data(out_r, np.int32).cpu().numpy()
    (a, av), (k, kv), (b, bv) = host(res[0], np.int32), host(res[1], np.int64), host(res[2], np.float64)
    rows = [(int(a[i]) if av[i] else None, int(k[i]) if kv[i] else None, float(b[i]) if bv[i] else None, int(li[i]), int(ri[i]))
            for i in range(n)]
    for c in res + [out_l, out_r]:
        api.gdf_column_free(c)
    return rows


@pytest.mark.parametrize("kind", ["inner", "left", "full"])
def test_result_cols_left_full_with_masked_payloads(ref, kind):
    """The part of the gather that is easy to get wrong (reference gdf_table.cuh:1232-1240,168-214): output rows whose
    index is -1 get validity bit 0 and their data is undefined; a payload's own NULLs travel with the row.  The product,
    the reference build and the expectation computed from the join indices agree row for row (as multisets)."""
    nl, nr = 4000, 2500
    lk, rk = np.random.randint(0, 3000, nl).astype(np.int64), np.random.randint(0, 3000, nr).astype(np.int64)
    lp, rp = (np.arange(nl, dtype=np.int32) * 10), (np.arange(nr, dtype=np.float64) + 0.5)
    lpv, rpv = ref_mask(nl, 0.6), ref_mask(nr, 0.6)
    want_rows = _join_result_cols(ref, kind, lk, lp, lpv, rk, rp, rpv)
    got_rows = _join_result_cols(libgdf, kind, lk, lp, lpv, rk, rp, rpv)
    key = lambda t: tuple((x is None, x if x is not None else 0) for x in t)
    assert sorted(got_rows, key=key) == sorted(want_rows, key=key)
    # and both equal what the indices imply
    lbits, rbits = np.unpackbits(lpv, bitorder="little")[:nl].astype(bool), np.unpackbits(rpv, bitorder="little")[:nr].astype(bool)
    for a, k, b, li, ri in got_rows[:: max(1, len(got_rows) // 500)]:
        assert a == (int(lp[li]) if li >= 0 and lbits[li] else None)
        assert b == (float(rp[ri]) if ri >= 0 and rbits[ri] else None)
        if li >= 0:
            assert k == int(lk[li])
    if kind != "inner":
        assert any(r[4] == -1 for r in got_rows)
    if kind == "full":
        assert any(r[3] == -1 for r in got_rows)
