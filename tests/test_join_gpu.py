"""-m gpu: hash joins through the C ABI vs the C oracle.
Case list follows the reference's gtest (src/tests/join/join-tests.cu:597-714: EqualValues 100x1000
range 1, MaxRandomValues 10k x 10k, Left/RightColumnsBigger, EmptyLeft/Right/Both) over 1-5 key
columns of int32/int64/float/double, INNER/LEFT/FULL (:516-572), every HASH case with a validity mask
whose second half is a coin flip (:216-220, valid_vectors.h:32-50).  Both results are sorted before
comparison, as the reference does (:341-345,464-467)."""
import numpy as np
import pytest
import torch

import oracle
from libgdf_b200 import columns as C
from libgdf_b200.libgdf_cffi import GDFError, ffi, libgdf
from gpu_utils import join, sorted_pairs

pytestmark = pytest.mark.gpu
KIND = {"inner": oracle.JOIN_INNER, "left": oracle.JOIN_LEFT, "full": oracle.JOIN_FULL}


def ref_style_mask(n):
    bits = np.ones(n, dtype=bool)
    bits[n // 2:] = np.random.rand(n - n // 2) < 0.5
    return np.packbits(bits, bitorder="little")


def check(kind, lcols, rcols, lvalid=None, rvalid=None):
    gl, gr = join(kind, lcols, rcols, lvalid, rvalid)
    ol, orr = oracle.join(KIND[kind], lcols, rcols, lvalid, rvalid)
    assert len(gl) == len(ol), "pair count %d vs %d" % (len(gl), len(ol))
    np.testing.assert_array_equal(sorted_pairs(gl, gr), sorted_pairs(ol, orr))


def make_tables(key_types, nl, nr, lrange, rrange=None, masks=True):
    rrange = rrange or lrange
    l = [np.random.randint(0, lrange, nl).astype(t) for t in key_types]
    r = [np.random.randint(0, rrange, nr).astype(t) for t in key_types]
    lv = [ref_style_mask(nl) for _ in key_types] if masks else None
    rv = [ref_style_mask(nr) for _ in key_types] if masks else None
    return l, r, lv, rv


KEYSETS = [[np.int32], [np.int64], [np.float32], [np.float64], [np.int32, np.int64], [np.int64, np.int32],
           [np.int32, np.float64, np.int64], [np.int8, np.int16, np.int32, np.int64, np.float64]]


@pytest.mark.parametrize("kind", ["inner", "left", "full"])
@pytest.mark.parametrize("key_types", KEYSETS, ids=lambda ts: "-".join(np.dtype(t).name for t in ts))
def test_equal_values(key_types, kind):
    l, r, lv, rv = make_tables(key_types, 100, 1000, 1)
    check(kind, l, r, lv, rv)


@pytest.mark.parametrize("kind", ["inner", "left", "full"])
@pytest.mark.parametrize("key_types", KEYSETS, ids=lambda ts: "-".join(np.dtype(t).name for t in ts))
def test_max_random_values(key_types, kind):
    l, r, lv, rv = make_tables(key_types, 10_000, 10_000, 2000 if len(key_types) == 1 else 12)
    check(kind, l, r, lv, rv)


@pytest.mark.parametrize("kind", ["inner", "left", "full"])
@pytest.mark.parametrize("shape", [(10_000, 100), (100, 10_000)])
def test_one_side_bigger(shape, kind):
    l, r, lv, rv = make_tables([np.int64, np.int32], shape[0], shape[1], 50, 50)
    check(kind, l, r, lv, rv)
    l, r, lv, rv = make_tables([np.int64], shape[0], shape[1], 5000, 5000, masks=False)
    check(kind, l, r)


@pytest.mark.parametrize("kind", ["inner", "left", "full"])
@pytest.mark.parametrize("sizes", [(0, 100), (100, 0), (0, 0)])
def test_empty_sides(sizes, kind):
    nl, nr = sizes
    l = [np.random.randint(0, 10, nl).astype(np.int32)]
    r = [np.random.randint(0, 10, nr).astype(np.int32)]
    if kind == "full" and nl == 0 and nr == 0:
        # reference: both-empty returns GDF_SUCCESS before trivial_full_join (joining.cu:303-305)
        gl, gr = join(kind, l, r)
        assert len(gl) == 0
        return
    gl, gr = join(kind, l, r)
    ol, orr = oracle.join(KIND[kind], l, r)
    np.testing.assert_array_equal(sorted_pairs(gl, gr), sorted_pairs(ol, orr))


@pytest.mark.parametrize("kind", ["inner", "left", "full"])
@pytest.mark.parametrize("key_t", [np.int64, np.int32])
def test_partitioned_path_unique_build(key_t, kind):
    """Build side above the partitioning threshold (2^20 rows): exercises join_part.cu with NP > 1.
    C3 shape at test size: build = permutation (unique), probe uniform over twice the range (50 % hit)."""
    nb, npr = 3_000_000, 5_000_000
    build = np.random.permutation(nb).astype(key_t)
    probe = np.random.randint(0, 2 * nb, npr).astype(key_t)
    check(kind, [probe], [build])


@pytest.mark.parametrize("kind", ["inner", "left", "full"])
def test_partitioned_path_duplicates_and_nulls(kind):
    nb, npr = 2_500_000, 3_000_000
    build = np.random.randint(0, 1_000_000, nb).astype(np.int64)     # ~2.5 duplicates per key
    probe = np.random.randint(0, 1_200_000, npr).astype(np.int64)
    lv = [np.packbits(np.random.rand(npr) < 0.7, bitorder="little")]
    rv = [np.packbits(np.random.rand(nb) < 0.7, bitorder="little")]
    check(kind, [probe], [build], lv, rv)


def test_inner_join_flips_to_build_on_smaller_side():
    left = np.random.permutation(2_200_000).astype(np.int64)          # left smaller -> built on left
    right = np.random.randint(0, 2_200_000, 4_000_000).astype(np.int64)
    check("inner", [left], [right])


def test_key_equal_to_empty_marker_falls_back():
    nb = 1_500_000
    build = np.random.permutation(nb).astype(np.int64)
    build[7] = -1
    probe = np.random.randint(-1, nb, 2_000_000).astype(np.int64)
    check("inner", [probe], [build])
    check("left", [probe], [build])


def test_c5_shape_composite_key_with_nulls():
    """BASELINE config C5 at test size: (int64,int32) key, 30 % null rows, left join."""
    nl, nr = 500_000, 50_000
    l = [np.random.randint(0, nr, nl).astype(np.int64), np.random.randint(0, 4, nl).astype(np.int32)]
    r = [np.random.randint(0, nr, nr).astype(np.int64), np.random.randint(0, 4, nr).astype(np.int32)]

    def masks(n):
        null_rows = np.random.rand(n) < 0.3
        which = np.random.rand(n) < 0.5
        return [np.packbits(~(null_rows & which), bitorder="little"), np.packbits(~(null_rows & ~which), bitorder="little")]
    check("left", l, r, masks(nl), masks(nr))


def test_result_cols_materialisation():
    """gdf_inner_join with result_cols: layout [left non-key.., key.., right non-key..]
    (reference joining.cu:412-439)."""
    nl, nr = 5000, 3000
    lk, lp = np.random.randint(0, 2000, nl).astype(np.int64), np.arange(nl, dtype=np.int32) * 10
    rk, rp = np.random.randint(0, 2000, nr).astype(np.int64), np.arange(nr, dtype=np.float64) + 0.5
    L = [C.column(lp), C.column(lk)]
    R = [C.column(rk), C.column(rp)]
    res = [ffi.new("gdf_column*") for _ in range(3)]
    res_arr = ffi.new("gdf_column*[]", res)
    ctx = ffi.new("gdf_context*")
    libgdf.gdf_context_view(ctx, 0, libgdf.GDF_HASH, 0, 0, 0)
    out_l, out_r = ffi.new("gdf_column*"), ffi.new("gdf_column*")
    libgdf.gdf_inner_join(C.column_array(L), 2, ffi.new("int[]", [1]), C.column_array(R), 2, ffi.new("int[]", [0]),
                          1, 3, res_arr, out_l, out_r, ctx)
    torch.cuda.synchronize()
    n = int(out_l.size)
    li = C.alias_column_data(out_l, np.int32).cpu().numpy()
    ri = C.alias_column_data(out_r, np.int32).cpu().numpy()
    ol, orr = oracle.join(oracle.JOIN_INNER, [lk], [rk])
    np.testing.assert_array_equal(sorted_pairs(li, ri), sorted_pairs(ol, orr))
    got_lp = C.alias_column_data(res[0], np.int32).cpu().numpy()
    got_key = C.alias_column_data(res[1], np.int64).cpu().numpy()
    got_rp = C.alias_column_data(res[2], np.float64).cpu().numpy()
    np.testing.assert_array_equal(got_lp, lp[li])
    np.testing.assert_array_equal(got_key, lk[li])
    np.testing.assert_array_equal(got_rp, rp[ri])
    assert int(res[0].size) == n and res[1].dtype == libgdf.GDF_INT64
    for c in res + [out_l, out_r]:
        libgdf.gdf_column_free(c)


def test_error_codes():
    ctx = ffi.new("gdf_context*")
    libgdf.gdf_context_view(ctx, 0, libgdf.GDF_HASH, 0, 0, 0)
    a, b = C.column(np.zeros(4, np.int32)), C.column(np.zeros(4, np.int64))
    o1, o2 = ffi.new("gdf_column*"), ffi.new("gdf_column*")
    idx = ffi.new("int[]", [0])
    with pytest.raises(GDFError) as e:
        libgdf.gdf_inner_join(C.column_array([a]), 1, idx, C.column_array([b]), 1, idx, 1, 0, ffi.NULL, o1, o2, ctx)
    assert e.value.errcode == "GDF_JOIN_DTYPE_MISMATCH"
    with pytest.raises(GDFError) as e:
        libgdf.gdf_inner_join(C.column_array([a]), 1, idx, C.column_array([a]), 1, idx, 1, 0, ffi.NULL, o1, o2, ffi.NULL)
    assert e.value.errcode == "GDF_INVALID_API_CALL"
    with pytest.raises(GDFError) as e:
        libgdf.gdf_inner_join(C.column_array([a]), 1, idx, C.column_array([a]), 1, idx, 1, 0, ffi.NULL, ffi.NULL, ffi.NULL, ctx)
    assert e.value.errcode == "GDF_DATASET_EMPTY"


@pytest.mark.parametrize("kind", ["inner", "left", "full"])
@pytest.mark.parametrize("key_types", [(np.int64, np.int32), (np.int32, np.int32)], ids=["i64-i32", "i32-i32"])
def test_partitioned_path_composite_key_with_nulls(key_types, kind):
    """(4/8-byte integer, 4-byte integer) composite keys take the radix-partitioned path with the second key in
    the slot's spare word (C5's key shape).  Build side above the partitioning threshold, duplicate composite
    keys, 30 % null rows with the null in either column."""
    nl, nr = 3_000_000, 1_600_000
    l = [np.random.randint(0, 900_000, nl).astype(key_types[0]), np.random.randint(0, 4, nl).astype(key_types[1])]
    r = [np.random.randint(0, 900_000, nr).astype(key_types[0]), np.random.randint(0, 4, nr).astype(key_types[1])]

    def masks(n):
        null_rows, which = np.random.rand(n) < 0.3, np.random.rand(n) < 0.5
        return [np.packbits(~(null_rows & which), bitorder="little"), np.packbits(~(null_rows & ~which), bitorder="little")]
    check(kind, l, r, masks(nl), masks(nr))


def test_partitioned_path_composite_key_same_first_key_different_second():
    """Rows that agree on the first key column but not on the second must not match."""
    nb = 1_300_000
    b0 = np.random.permutation(nb).astype(np.int64)
    build = [b0, np.zeros(nb, np.int32)]
    probe = [b0[:2_000_000 % nb or nb].copy(), np.ones(min(2_000_000 % nb or nb, nb), np.int32)]
    probe[1][::3] = 0                                              # every third probe row matches
    check("inner", probe, build)
    check("left", probe, build)


# ---- compact (32-bit key) path of the partitioned join: csrc/join_compact.cuh ----
@pytest.mark.parametrize("kind", ["inner", "left", "full"])
@pytest.mark.parametrize("hit", [0.0, 0.03, 1.0], ids=["nohit", "hit3pct", "allhit"])
def test_compact_path_hit_rates(kind, hit):
    """Unique build keys that fit 32 bits: INNER fills per-warp output chunks and the fix-up pass closes the holes
    (0 % and 3 % hit rates leave almost every chunk nearly empty, 100 % fills them), LEFT/FULL write in place."""
    nb, npr = 2_300_000, 4_100_003
    build = (np.random.permutation(nb).astype(np.int64) * 3 + 1)
    if hit == 0.0:
        probe = np.random.randint(0, nb, npr).astype(np.int64) * 3               # never 1 mod 3
    elif hit == 1.0:
        probe = build[np.random.randint(0, nb, npr)]
    else:
        probe = np.random.randint(0, int(3 * nb / hit), npr).astype(np.int64)
    check(kind, [probe], [build])


@pytest.mark.parametrize("kind", ["inner", "left", "full"])
def test_compact_path_probe_keys_wider_than_32_bits(kind):
    """Build keys fit 32 bits, some probe keys do not (and some are negative): those rows can never match - INNER
    drops them, LEFT/FULL emit (row,-1) - and must not alias a build key with the same low word."""
    nb, npr = 1_400_000, 2_000_000
    build = np.random.permutation(nb).astype(np.int64)
    probe = np.random.randint(0, nb, npr).astype(np.int64)
    probe[::5] += np.int64(1) << 32                    # same low word as a build key, different high word
    probe[1::7] = -probe[1::7] - 1                     # negative: high word all ones
    check(kind, [probe], [build])


@pytest.mark.parametrize("kind", ["inner", "left"])
def test_wide_build_keys_take_the_general_path(kind):
    nb, npr = 1_400_000, 2_000_000
    build = np.random.permutation(nb).astype(np.int64) + (np.int64(5) << 32)
    build[::2] -= np.int64(5) << 32
    probe = np.concatenate([build[np.random.randint(0, nb, npr // 2)], np.random.randint(0, nb, npr - npr // 2).astype(np.int64)])
    check(kind, [probe], [build])


@pytest.mark.parametrize("kind", ["inner", "left", "full"])
def test_compact_path_int32_negative_keys_and_masks(kind):
    nb, npr = 1_300_000, 1_900_000
    build = (np.random.permutation(nb).astype(np.int32) - nb // 2)
    probe = np.random.randint(-nb, nb, npr).astype(np.int32)
    lv = [np.packbits(np.random.rand(npr) < 0.8, bitorder="little")]
    rv = [np.packbits(np.random.rand(nb) < 0.9, bitorder="little")]
    check(kind, [probe], [build], lv, rv)


def test_compact_path_heavy_duplicates():
    """A few build keys repeated thousands of times: chains span many buckets; exact count pass + write pass."""
    nb, npr = 1_200_000, 6_000
    build = np.random.randint(0, 300, nb).astype(np.int64)
    probe = np.random.randint(0, 400, npr).astype(np.int64)
    gl, gr = join("inner", [probe], [build])
    counts = np.bincount(build, minlength=400)
    assert len(gl) == int(counts[probe].sum())
    assert np.array_equal(probe[gl], build[gr])
    pairs = gl.astype(np.int64) * nb + gr
    assert len(np.unique(pairs)) == len(pairs)


# ---- histogram-free probe side of the compact INNER path (fixed partition regions + overflow fallback) ----
@pytest.mark.parametrize("np_rows", [1_048_576, 3_000_017, 5_250_000])
def test_padded_probe_layout_uniform_keys(np_rows):
    """INNER, unique 32-bit build keys, >= 2^20 probe rows: the probe side is partitioned without a histogram pass into
    fixed regions; region ends, partial and empty tiles are exercised by sizes that do not divide anything."""
    nb = 1_300_000
    build = np.random.permutation(nb).astype(np.int64)
    probe = np.random.randint(0, 2 * nb, np_rows).astype(np.int64)
    check("inner", [probe], [build])


@pytest.mark.parametrize("hot", [0.2, 0.9])
def test_padded_probe_layout_overflows_on_skewed_keys_and_falls_back(hot):
    """A hot probe key overflows its partition's region: the scatter drops what does not fit (raising a flag, never
    writing out of bounds), the partial result is discarded and the side is partitioned again with exact counts."""
    nb, npr = 2_200_000, 4_000_000
    build = np.random.permutation(nb).astype(np.int64)
    probe = np.random.randint(0, nb, npr).astype(np.int64)
    probe[np.random.rand(npr) < hot] = 12345
    check("inner", [probe], [build])
    check("inner", [build], [probe[:1_100_000]])      # flipped sizes: the hot key is now on the build side (duplicates -> exact path)
