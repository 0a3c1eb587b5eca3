"""CPU: argument validation of the hot-path entry points happens on the host, before any CUDA call, and returns
the reference's error codes (SURVEY.md section 8b "Error convention").  Columns carry fake, never dereferenced
device addresses; every call here must fail (or succeed trivially) without touching a GPU."""
import pytest

FAKE = 0x10000   # never dereferenced: validation rejects the call first


@pytest.fixture()
def h(gdf):
    ffi, libgdf = gdf

    class H(object):
        pass
    o = H()
    o.ffi, o.lib = ffi, libgdf

    def col(size, dtype, valid=False, data=True):
        c = ffi.new("gdf_column*")
        libgdf.gdf_column_view(c, ffi.cast("void*", FAKE) if data else ffi.NULL,
                               ffi.cast("gdf_valid_type*", FAKE) if valid else ffi.NULL, size, getattr(libgdf, dtype))
        return c
    o.col = col
    o.arr = lambda cols: ffi.new("gdf_column*[]", cols)
    o.ctx = ffi.new("gdf_context*")
    libgdf.gdf_context_view(o.ctx, 0, libgdf.GDF_HASH, 0, 0, 0)
    return o


def code(h, fn, *args):
    from libgdf_b200.libgdf_cffi import GDFError
    try:
        fn(*args)
    except GDFError as e:
        return e.errcode
    return "GDF_SUCCESS"


def join_args(h, l, r, ctx=None, n=None):
    n = n or len(l)
    idx = h.ffi.new("int[]", list(range(n)))
    o1, o2 = h.ffi.new("gdf_column*"), h.ffi.new("gdf_column*")
    return (h.arr(l), len(l), idx, h.arr(r), len(r), idx, n, 0, h.ffi.NULL, o1, o2, ctx if ctx is not None else h.ctx)


@pytest.mark.parametrize("fn", ["gdf_inner_join", "gdf_left_join", "gdf_full_join"])
def test_join_validation(h, fn):
    f = getattr(h.lib, fn)
    a64, b64, a32 = h.col(10, "GDF_INT64"), h.col(20, "GDF_INT64"), h.col(10, "GDF_INT32")
    # reference joining.cu:300-301: each side must have fewer than INT_MAX rows
    big = h.col(2 ** 31 - 1, "GDF_INT64")
    assert code(h, f, *join_args(h, [big], [b64])) == "GDF_COLUMN_SIZE_TOO_BIG"
    assert code(h, f, *join_args(h, [a64], [a32])) == "GDF_JOIN_DTYPE_MISMATCH"            # :334
    assert code(h, f, *join_args(h, [a64, b64], [a64, a64])) == "GDF_COLUMN_SIZE_MISMATCH"  # :335-336
    assert code(h, f, *join_args(h, [a64], [b64], ctx=h.ffi.NULL)) == "GDF_INVALID_API_CALL"  # :293-294
    sort_ctx = h.ffi.new("gdf_context*")
    h.lib.gdf_context_view(sort_ctx, 0, h.lib.GDF_SORT, 0, 0, 0)
    assert code(h, f, *join_args(h, [a64, a64], [b64, b64], ctx=sort_ctx)) == "GDF_JOIN_TOO_MANY_COLUMNS"  # :352-367
    assert code(h, f, *join_args(h, [a64], [b64], ctx=sort_ctx)) == "GDF_UNSUPPORTED_METHOD"   # sort-merge join: out of scope
    # no data pointer on a non-empty side (:327-331)
    assert code(h, f, *join_args(h, [h.col(10, "GDF_INT64", data=False)], [b64])) == "GDF_DATASET_EMPTY"
    # both sides empty: success, outputs untouched (:303-305)
    e1, e2 = h.col(0, "GDF_INT64", data=False), h.col(0, "GDF_INT64", data=False)
    assert code(h, f, *join_args(h, [e1], [e2])) == "GDF_SUCCESS"


def test_inner_and_left_join_empty_side_short_cuts(h):
    a, e = h.col(10, "GDF_INT64"), h.col(0, "GDF_INT64", data=False)
    assert code(h, h.lib.gdf_inner_join, *join_args(h, [a], [e])) == "GDF_SUCCESS"     # joining.cu:311-317
    assert code(h, h.lib.gdf_inner_join, *join_args(h, [e], [a])) == "GDF_SUCCESS"
    assert code(h, h.lib.gdf_left_join, *join_args(h, [e], [a])) == "GDF_SUCCESS"      # :307-309


@pytest.mark.parametrize("fn", ["gdf_group_by_sum", "gdf_group_by_min", "gdf_group_by_max", "gdf_group_by_avg",
                                "gdf_group_by_count"])
def test_groupby_validation(h, fn):
    f = getattr(h.lib, fn)
    k, v, ok, ov = h.col(8, "GDF_INT64"), h.col(8, "GDF_INT64"), h.col(8, "GDF_INT64"), h.col(8, "GDF_INT64")
    # any validity mask is rejected (sqls_ops.cu:1103-1106)
    assert code(h, f, 1, h.arr([h.col(8, "GDF_INT64", valid=True)]), v, h.ffi.NULL, h.arr([ok]), ov, h.ctx) == "GDF_VALIDITY_UNSUPPORTED"
    assert code(h, f, 1, h.arr([k]), h.col(8, "GDF_INT64", valid=True), h.ffi.NULL, h.arr([ok]), ov, h.ctx) == "GDF_VALIDITY_UNSUPPORTED"
    # null argument set (:1095-1102)
    assert code(h, f, 1, h.arr([k]), h.ffi.NULL, h.ffi.NULL, h.arr([ok]), ov, h.ctx) == "GDF_DATASET_EMPTY"
    assert code(h, f, 1, h.arr([k]), v, h.ffi.NULL, h.arr([ok]), ov, h.ffi.NULL) == "GDF_DATASET_EMPTY"
    # empty input: sizes set to 0, success (:1110-1128)
    e, ev, eo, eov = h.col(0, "GDF_INT64", data=False), h.col(0, "GDF_INT64", data=False), h.col(5, "GDF_INT64"), h.col(5, "GDF_INT64")
    assert code(h, f, 1, h.arr([e]), ev, h.ffi.NULL, h.arr([eo]), eov, h.ctx) == "GDF_SUCCESS"
    assert eo.size == 0 and eov.size == 0
    # GDF_SORT and flag_sort_result are implemented since round 2 (tests/test_sort_gpu.py): with valid arguments they are
    # no longer refused on the host; a method outside the enum still is (sqls_ops.cu:1086-1093)
    bad_ctx = h.ffi.new("gdf_context*")
    h.lib.gdf_context_view(bad_ctx, 0, h.lib.GDF_HASH, 0, 0, 0)
    bad_ctx.flag_method = h.lib.N_GDF_METHODS
    assert code(h, f, 1, h.arr([k]), v, h.ffi.NULL, h.arr([ok]), ov, bad_ctx) == "GDF_UNSUPPORTED_METHOD"


def test_filter_comparison_stencil_validation(h):
    lib, ffi = h.lib, h.ffi
    # gdf_filter: a mask on the first column is rejected before anything else (sqls_ops.cu:1412)
    cols = ffi.new("gdf_column[]", 1)
    cols[0] = h.col(16, "GDF_INT32", valid=True)[0]
    sz = ffi.new("size_t*")
    assert code(h, lib.gdf_filter, 16, cols, 1, ffi.NULL, ffi.NULL, ffi.NULL, ffi.NULL, sz) == "GDF_VALIDITY_UNSUPPORTED"
    cols[0] = h.col(16, "GDF_DATE64")[0]                       # dtype outside INT8..FLOAT64 (sqls_rtti_comp.hpp:200-213)
    assert code(h, lib.gdf_filter, 16, cols, 1, ffi.NULL, ffi.NULL, ffi.NULL, ffi.NULL, sz) == "GDF_UNSUPPORTED_DTYPE"
    # gpu_comparison / _static: size mismatch and a non-int8 output both report COLUMN_SIZE_MISMATCH (filterops.cu:163-165,264)
    a, b, o8, o32 = h.col(8, "GDF_INT64"), h.col(9, "GDF_INT64"), h.col(8, "GDF_INT8", valid=True), h.col(8, "GDF_INT32", valid=True)
    assert code(h, lib.gpu_comparison, a, b, o8, lib.GDF_EQUALS) == "GDF_COLUMN_SIZE_MISMATCH"
    assert code(h, lib.gpu_comparison, a, a, o32, lib.GDF_EQUALS) == "GDF_COLUMN_SIZE_MISMATCH"
    assert code(h, lib.gpu_comparison_static_i64, a, 3, h.col(9, "GDF_INT8", valid=True), lib.GDF_EQUALS) == "GDF_COLUMN_SIZE_MISMATCH"
    assert code(h, lib.gpu_comparison_static_i64, a, 3, o32, lib.GDF_EQUALS) == "GDF_COLUMN_SIZE_MISMATCH"
    # gpu_apply_stencil (streamcompactionops.cu:208-219)
    st = h.col(8, "GDF_INT8", valid=True)
    assert code(h, lib.gpu_apply_stencil, h.col(8, "GDF_INT64", valid=True), st, h.col(8, "GDF_INT64", valid=True)) == "GDF_VALIDITY_UNSUPPORTED"
    assert code(h, lib.gpu_apply_stencil, a, st, h.col(8, "GDF_INT32", valid=True)) == "GDF_DTYPE_MISMATCH"
    assert code(h, lib.gpu_apply_stencil, a, st, h.col(7, "GDF_INT64", valid=True)) == "GDF_COLUMN_SIZE_MISMATCH"


def test_binary_ops_hash_partition_validation(h):
    lib, ffi = h.lib, h.ffi
    a, b, o = h.col(8, "GDF_INT32"), h.col(9, "GDF_INT32"), h.col(8, "GDF_INT32")
    assert code(h, lib.gdf_add_generic, a, b, o) == "GDF_COLUMN_SIZE_MISMATCH"              # binaryops.cu:42-44
    assert code(h, lib.gdf_add_generic, a, h.col(8, "GDF_INT64"), o) == "GDF_UNSUPPORTED_DTYPE"   # :45
    assert code(h, lib.gdf_add_i32, h.col(8, "GDF_INT64"), h.col(8, "GDF_INT64"), o) == "GDF_UNSUPPORTED_DTYPE"
    assert code(h, lib.gdf_add_generic, h.col(0, "GDF_INT32", data=False), h.col(0, "GDF_INT32", data=False), o) == "GDF_SUCCESS"  # :38-41
    # gdf_hash (hashing.cu:83-110)
    assert code(h, lib.gdf_hash, 1, h.arr([a]), lib.GDF_HASH_MURMUR3, h.col(8, "GDF_INT64")) == "GDF_UNSUPPORTED_DTYPE"
    assert code(h, lib.gdf_hash, 1, h.arr([a]), 7, h.col(8, "GDF_INT32")) == "GDF_INVALID_HASH_FUNCTION"
    assert code(h, lib.gdf_hash, 0, h.arr([a]), lib.GDF_HASH_MURMUR3, h.col(8, "GDF_INT32")) == "GDF_DATASET_EMPTY"
    # gdf_hash_partition (hashing.cu:573-607)
    offs, idx = ffi.new("int[]", 4), ffi.new("int[]", [0])
    assert code(h, lib.gdf_hash_partition, 1, h.arr([a]), idx, 1, 4, h.arr([h.col(8, "GDF_INT64")]), offs, lib.GDF_HASH_MURMUR3) == "GDF_PARTITION_DTYPE_MISMATCH"
    assert code(h, lib.gdf_hash_partition, 1, h.arr([a]), idx, 1, 4, h.arr([h.col(9, "GDF_INT32")]), offs, lib.GDF_HASH_MURMUR3) == "GDF_COLUMN_SIZE_MISMATCH"
    assert code(h, lib.gdf_hash_partition, 1, h.arr([a]), idx, 1, 4, h.arr([o]), offs, 7) == "GDF_INVALID_HASH_FUNCTION"
    assert code(h, lib.gdf_hash_partition, 1, h.arr([a]), idx, 1, 0, h.arr([o]), offs, lib.GDF_HASH_MURMUR3) == "GDF_INVALID_API_CALL"
