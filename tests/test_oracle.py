"""CPU tests: the oracle against the reference's golden vectors and known-answer hashes.
(No GPU, no product code: this pins the checker itself.)"""
import json
import os
import struct

import numpy as np
import pytest

import oracle
from oracle import np_oracle

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.json")))


def _py_murmur3_32(data: bytes, seed=0):
    """Independent pure-python MurmurHash3_x86_32 (public algorithm), small inputs only."""
    c1, c2, m = 0xcc9e2d51, 0x1b873593, 0xffffffff
    h = seed
    rounded = len(data) & ~3
    for i in range(0, rounded, 4):
        k = struct.unpack_from("<I", data, i)[0]
        k = (k * c1) & m
        k = ((k << 15) | (k >> 17)) & m
        k = (k * c2) & m
        h ^= k
        h = ((h << 13) | (h >> 19)) & m
        h = (h * 5 + 0xe6546b64) & m
    k = 0
    tail = data[rounded:]
    if len(tail) >= 3: k ^= tail[2] << 16
    if len(tail) >= 2: k ^= tail[1] << 8
    if len(tail) >= 1:
        k ^= tail[0]
        k = (k * c1) & m
        k = ((k << 15) | (k >> 17)) & m
        k = (k * c2) & m
        h ^= k
    h ^= len(data)
    h ^= h >> 16
    h = (h * 0x85ebca6b) & m
    h ^= h >> 13
    h = (h * 0xc2b2ae35) & m
    h ^= h >> 16
    return h


def test_murmur_known_answers():
    kat = GOLDEN["murmur3_32"]
    for text, want in kat["bytes"].items():
        assert oracle.murmur3_32(text.encode()) == int(want, 16) == _py_murmur3_32(text.encode())
    for np_t in ("int64", "int32", "int16", "int8"):
        for value, want in kat[np_t].items():
            raw = np.dtype(np_t).type(int(value)).tobytes()
            assert oracle.murmur3_32(raw) == int(want, 16), (np_t, value)
            assert _py_murmur3_32(raw) == int(want, 16)


def test_row_hash_combine_known_answer():
    kat = GOLDEN["murmur3_32"]["row_int64_int32"]
    k0 = np.array([kat["key0"]], dtype=np.int64)
    k1 = np.array([kat["key1"]], dtype=np.int32)
    assert oracle.hash_rows([k0])[0].view(np.uint32) == int(kat["hash0"], 16)
    assert oracle.hash_rows([k1])[0].view(np.uint32) == int(kat["hash1"], 16)
    assert oracle.hash_rows([k0, k1])[0].view(np.uint32) == int(kat["combined"], 16)


def test_hash_random_against_python():
    rng = np.random.RandomState(7)
    for np_t in (np.int8, np.int16, np.int32, np.int64, np.float32, np.float64):
        col = (rng.randn(50) * 1000).astype(np_t)
        got = oracle.hash_rows([col]).view(np.uint32)
        want = [_py_murmur3_32(col[i].tobytes()) for i in range(len(col))]
        assert got.tolist() == want


def test_identity_hash():
    col = np.array([-1, 0, 7, 2 ** 40 + 5], dtype=np.int64)
    assert oracle.hash_rows([col], identity=True).view(np.uint32).tolist() == [0xffffffff, 0, 7, 5]
    col8 = np.array([-1, 3], dtype=np.int8)
    assert oracle.hash_rows([col8], identity=True).view(np.uint32).tolist() == [0xffffffff, 3]


def test_filter_golden():
    g = GOLDEN["filter"]
    cols = [np.array(c, dtype=t) for c, t in zip(g["cols"], g["dtypes"])]
    assert np_oracle.filter_rows(cols, g["vals"]).tolist() == g["indices"]
    assert oracle.filter_i64(np.array([3, 1, 3, 3, 0]), 3).tolist() == [0, 2, 3]


@pytest.mark.parametrize("name,op,agg_key", [("sum", oracle.OP_SUM, "agg_sum_count_avg"),
                                             ("avg", oracle.OP_AVG, "agg_sum_count_avg"),
                                             ("min", oracle.OP_MIN, "agg_min_max"),
                                             ("max", oracle.OP_MAX, "agg_min_max")])
def test_groupby_golden(name, op, agg_key):
    g = GOLDEN["groupby"]
    keys = [np.array(c, dtype=t) for c, t in zip(g["keys"], g["key_dtypes"])]
    vals = np.array(g[agg_key], dtype=np.float64)
    out_keys, out_agg = oracle.groupby(op, keys, vals, out_dtype=oracle.GDF_FLOAT64)
    got = sorted(zip(*[k.tolist() for k in out_keys], out_agg.tolist()))
    want = sorted(zip(*g["group_keys_sorted"], g[name]))
    assert got == want


def test_groupby_count_golden():
    g = GOLDEN["groupby"]
    keys = [np.array(c, dtype=t) for c, t in zip(g["keys"], g["key_dtypes"])]
    vals = np.array(g["agg_sum_count_avg"], dtype=np.float64)
    out_keys, out_agg = oracle.groupby(oracle.OP_COUNT, keys, vals, out_dtype=oracle.GDF_INT32)
    got = sorted(zip(*[k.tolist() for k in out_keys], out_agg.tolist()))
    assert got == sorted(zip(*g["group_keys_sorted"], g["count"]))


def test_groupby_int8_sum_wraps():
    keys = [np.zeros(300, dtype=np.int32)]
    vals = np.ones(300, dtype=np.int8)
    _, agg = oracle.groupby(oracle.OP_SUM, keys, vals)
    assert agg.dtype == np.int8 and agg.tolist() == [np.int8(300 % 256 - 0 if 300 % 256 < 128 else 300 % 256 - 256)]


def test_join_against_bruteforce():
    # same oracle rule as the reference's test (src/tests/join/join-tests.cu:260-356): std::multimap on
    # the keys, nulls never match, LEFT/FULL pad with -1, results compared sorted.
    rng = np.random.RandomState(3)
    for kind in (oracle.JOIN_INNER, oracle.JOIN_LEFT, oracle.JOIN_FULL):
        l0, l1 = rng.randint(0, 6, 40).astype(np.int64), rng.randint(0, 3, 40).astype(np.int32)
        r0, r1 = rng.randint(0, 6, 25).astype(np.int64), rng.randint(0, 3, 25).astype(np.int32)
        lv = [np.packbits(rng.rand(40) > 0.3, bitorder="little"), None]
        rv = [None, np.packbits(rng.rand(25) > 0.3, bitorder="little")]
        gl, gr = oracle.join(kind, [l0, l1], [r0, r1], lv, rv)
        lval = np_oracle.unpack_valid(lv[0], 40)
        rval = np_oracle.unpack_valid(rv[1], 25)
        want = []
        matched_r = set()
        for i in range(40):
            hit = False
            if lval[i]:
                for j in range(25):
                    if rval[j] and l0[i] == r0[j] and l1[i] == r1[j]:
                        want.append((i, j)); matched_r.add(j); hit = True
            if not hit and kind != oracle.JOIN_INNER:
                want.append((i, -1))
        if kind == oracle.JOIN_FULL:
            want += [(-1, j) for j in range(25) if j not in matched_r]
        assert sorted(zip(gl.tolist(), gr.tolist())) == sorted(want)


def test_partition_ids():
    col = np.arange(1000, dtype=np.int64)
    h = oracle.hash_rows([col]).view(np.uint32)
    assert (oracle.partition_ids([col], 8) == (h & 7)).all()
    assert (oracle.partition_ids([col], 5) == (h % 5)).all()


def test_np_oracle_basics():
    a = np.array([1, 2, 3, 4, 5, 6, 7, 8, 9], dtype=np.int32)
    b = np.array([9, 8, 7, 6, 5, 4, 3, 2, 1], dtype=np.int32)
    mask = np.packbits(np.array([1, 0, 1, 1, 0, 1, 1, 1, 0], dtype=bool), bitorder="little")
    out = np_oracle.binary_op("add", a, b, np.full(9, -7, np.int32), lvalid=mask)
    assert out.tolist() == [10, -7, 10, 10, -7, 10, 10, 10, -7]
    assert np_oracle.reduce("sum", np.array([100, 100], dtype=np.int8)) == np.int8(-56)
    assert np_oracle.reduce("min", np.array([], dtype=np.int32)) == np.iinfo(np.int32).max
    data = np.arange(10, dtype=np.int64)
    st = np.array([1, 0, 1, 1, 0, 0, 1, 0, 1, 1], dtype=np.int8)
    kept, mask_out = np_oracle.apply_stencil(data, st, np.array([0xFF, 0xFF], dtype=np.uint8))
    assert kept.tolist() == [0, 2, 3, 6, 8, 9] and mask_out.tolist() == [0xFF, 0xC0]
    # MSB-first read of the stencil mask: byte 0b00000001 validates row 7 only
    kept, _ = np_oracle.apply_stencil(data[:8], np.ones(8, np.int8), np.array([0x01], dtype=np.uint8))
    assert kept.tolist() == [7]
