"""-m gpu: gdf_hash and gdf_hash_partition through the C ABI vs the C oracle.
Partition verification follows the reference's own test (src/tests/hashing/hash-partition-test.cu:166-254):
every output row must sit inside the [offset_p, offset_{p+1}) range of the partition its hash selects,
and the output must be a permutation of the input rows (with their validity bits)."""
import numpy as np
import pytest
import torch

import oracle
from oracle import np_oracle
from libgdf_b200 import columns as C
from libgdf_b200.libgdf_cffi import GDFError, ffi, libgdf
from gpu_utils import gen_rand, rand_mask

pytestmark = pytest.mark.gpu
NP_TYPES = [np.int8, np.int16, np.int32, np.int64, np.float32, np.float64]


def gdf_hash(cols_np, func="GDF_HASH_MURMUR3"):
    cols = [C.column(c) for c in cols_np]
    out = C.empty_column(len(cols_np[0]), torch.int32)
    libgdf.gdf_hash(len(cols), C.column_array(cols), getattr(libgdf, func), out.cdata)
    torch.cuda.synchronize()
    return out.to_numpy()


@pytest.mark.parametrize("np_t", NP_TYPES)
def test_hash_single_column_matches_oracle(np_t):
    col = gen_rand(np_t, 10_007)
    np.testing.assert_array_equal(gdf_hash([col]), oracle.hash_rows([col]))
    np.testing.assert_array_equal(gdf_hash([col], "GDF_HASH_IDENTITY"), oracle.hash_rows([col], identity=True))


def test_hash_multi_column_and_equal_rows():
    n = 50_000
    cols = [gen_rand(np.int64, n, 0, 50), gen_rand(np.int32, n, 0, 4), gen_rand(np.float64, n), gen_rand(np.int8, n)]
    np.testing.assert_array_equal(gdf_hash(cols), oracle.hash_rows(cols))
    # the reference's own property (hash-test.cu:35-158, test_hashing.py:61-85): equal rows hash equal
    dup = [np.concatenate([c, c]) for c in cols]
    h = gdf_hash(dup)
    np.testing.assert_array_equal(h[:n], h[n:])


def test_hash_errors():
    col = C.column(np.zeros(4, np.int32))
    out64 = C.empty_column(4, torch.int64)
    with pytest.raises(GDFError) as e:
        libgdf.gdf_hash(1, C.column_array([col]), libgdf.GDF_HASH_MURMUR3, out64.cdata)
    assert e.value.errcode == "GDF_UNSUPPORTED_DTYPE"


def run_partition(cols_np, hash_idx, nparts, func="GDF_HASH_MURMUR3", valids=None):
    n = len(cols_np[0])
    valids = valids or [None] * len(cols_np)
    cols = [C.column(c, v) for c, v in zip(cols_np, valids)]
    outs = [C.empty_column(n, c.data.dtype, with_valid=v is not None) for c, v in zip(cols, valids)]
    offsets = ffi.new("int[]", nparts)
    libgdf.gdf_hash_partition(len(cols), C.column_array(cols), ffi.new("int[]", hash_idx), len(hash_idx), nparts,
                              C.column_array(outs), offsets, getattr(libgdf, func))
    torch.cuda.synchronize()
    return [o.to_numpy() for o in outs], [None if o.valid is None else o.valid.cpu().numpy() for o in outs], list(offsets)


def check_partition(cols_np, hash_idx, nparts, outs, out_valids, offsets, identity=False, valids=None):
    n = len(cols_np[0])
    assert offsets[0] == 0 and all(a <= b for a, b in zip(offsets, offsets[1:])) and offsets[-1] <= n
    pid_out = oracle.partition_ids([outs[i] for i in hash_idx], nparts, identity=identity)
    bounds = offsets + [n]
    for p in range(nparts):
        seg = pid_out[bounds[p]:bounds[p + 1]]
        assert (seg == p).all(), "row outside its partition"
    # permutation check: the multiset of full rows (values + validity bits) is preserved
    def rows(cs, vs):
        parts = [c.astype(np.float64) if c.dtype.kind == "f" else c.astype(np.int64) for c in cs]
        for v in vs:
            if v is not None:
                parts.append(np_oracle.unpack_valid(v, n).astype(np.int64))
        return sorted(zip(*[p.tolist() for p in parts]))
    assert rows(outs, out_valids) == rows(cols_np, valids or [None] * len(cols_np))


@pytest.mark.parametrize("nparts", [1, 5, 8, 10, 257, 4096])
@pytest.mark.parametrize("n", [1, 1000, 100_003])
def test_partition_int64(n, nparts):
    cols = [gen_rand(np.int64, n, 0, 1 << 30), gen_rand(np.float64, n)]
    outs, ov, off = run_partition(cols, [0], nparts)
    check_partition(cols, [0], nparts, outs, ov, off)


def test_partition_multi_key_identity_and_masks():
    n = 60_001
    cols = [gen_rand(np.int32, n, 0, 1000), gen_rand(np.int16, n, 0, 100), gen_rand(np.int64, n)]
    valids = [rand_mask(n)[0], None, rand_mask(n)[0]]
    for func, ident in (("GDF_HASH_MURMUR3", False), ("GDF_HASH_IDENTITY", True)):
        outs, ov, off = run_partition(cols, [0, 1], 16, func, valids)
        check_partition(cols, [0, 1], 16, outs, ov, off, identity=ident, valids=valids)


def test_partition_errors():
    a = C.column(np.zeros(4, np.int32))
    o = C.empty_column(4, torch.int64)
    offsets = ffi.new("int[]", 2)
    with pytest.raises(GDFError) as e:
        libgdf.gdf_hash_partition(1, C.column_array([a]), ffi.new("int[]", [0]), 1, 2, C.column_array([o]), offsets,
                                  libgdf.GDF_HASH_MURMUR3)
    assert e.value.errcode == "GDF_PARTITION_DTYPE_MISMATCH"
