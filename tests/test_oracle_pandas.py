"""CPU: the C oracle cross-checked against pandas (an implementation written by someone else) on seeded random
tables - a second, independent pin of the checker besides the reference's golden vectors (tests/test_oracle.py)
and the reference's own kernels (tests/test_reference_parity.py, GPU)."""
import numpy as np
import pandas as pd
import pytest

import oracle

KIND = {"inner": oracle.JOIN_INNER, "left": oracle.JOIN_LEFT, "full": oracle.JOIN_FULL}
HOW = {"inner": "inner", "left": "left", "full": "outer"}


def _sorted(l, r):
    a = np.stack([np.asarray(l, dtype=np.int64), np.asarray(r, dtype=np.int64)], 1)
    return a[np.lexsort((a[:, 1], a[:, 0]))]


@pytest.mark.parametrize("kind", ["inner", "left", "full"])
@pytest.mark.parametrize("key_types", [(np.int64,), (np.int32,), (np.int64, np.int32), (np.int32, np.int8, np.int64)],
                         ids=lambda ts: "-".join(np.dtype(t).name for t in ts))
def test_join_matches_pandas_merge(kind, key_types):
    rng = np.random.RandomState(42)
    nl, nr = 4000, 1500
    span = 900 if len(key_types) == 1 else 9
    l = [rng.randint(0, span, nl).astype(t) for t in key_types]
    r = [rng.randint(0, span, nr).astype(t) for t in key_types]
    ol, orr = oracle.join(KIND[kind], l, r)
    cols = ["k%d" % i for i in range(len(key_types))]
    L = pd.DataFrame({c: v for c, v in zip(cols, l)})
    R = pd.DataFrame({c: v for c, v in zip(cols, r)})
    L["li"], R["ri"] = np.arange(nl), np.arange(nr)
    m = L.merge(R, on=cols, how=HOW[kind])
    pl = m["li"].fillna(-1).astype(np.int64).to_numpy()
    pr = m["ri"].fillna(-1).astype(np.int64).to_numpy()
    np.testing.assert_array_equal(_sorted(ol, orr), _sorted(pl, pr))


@pytest.mark.parametrize("op,name", [(oracle.OP_SUM, "sum"), (oracle.OP_MIN, "min"), (oracle.OP_MAX, "max"),
                                     (oracle.OP_COUNT, "count")])
@pytest.mark.parametrize("val_t", [np.int64, np.int32, np.float64])
def test_groupby_matches_pandas(op, name, val_t):
    rng = np.random.RandomState(7)
    n = 20_000
    k0, k1 = rng.randint(0, 40, n).astype(np.int64), rng.randint(0, 5, n).astype(np.int32)
    v = (rng.randint(-1000, 1000, n)).astype(val_t)
    gk, ga = oracle.groupby(op, [k0, k1], v)
    df = pd.DataFrame({"k0": k0, "k1": k1, "v": v}).groupby(["k0", "k1"])["v"].agg(name).reset_index()
    got = sorted(zip(gk[0].tolist(), gk[1].tolist(), ga.tolist()))
    want = sorted(zip(df["k0"].tolist(), df["k1"].tolist(), df["v"].astype(val_t if name != "count" else ga.dtype).tolist()))
    assert len(got) == len(want)
    for a, b in zip(got, want):
        assert a[:2] == b[:2]
        assert a[2] == pytest.approx(b[2], rel=1e-12, abs=1e-9)


def test_groupby_avg_matches_pandas_mean():
    rng = np.random.RandomState(11)
    n = 10_000
    k = rng.randint(0, 100, n).astype(np.int64)
    v = rng.rand(n)
    gk, ga = oracle.groupby(oracle.OP_AVG, [k], v)
    want = pd.Series(v).groupby(k).mean()
    got = dict(zip(gk[0].tolist(), ga.tolist()))
    assert set(got) == set(want.index.tolist())
    for key, mean in want.items():
        assert got[key] == pytest.approx(mean, rel=1e-12)


def test_filter_and_sum_match_numpy():
    rng = np.random.RandomState(3)
    col = rng.randint(0, 10, 100_003).astype(np.int64)
    np.testing.assert_array_equal(oracle.filter_i64(col, 3), np.nonzero(col == 3)[0].astype(np.uint64))
    big = rng.randint(-2 ** 62, 2 ** 62, 10_001).astype(np.int64)        # wraps like the column's own type
    assert oracle.sum_i64(big) == int(np.sum(big, dtype=np.int64))
