"""CPU (not gpu): the multi-GPU exchange plan of libgdf_b200/dist.py under the gloo backend with
world_size 2 and 3.  The per-shard operators are injected here from the CPU oracle (test infrastructure) -
what is under test is the host-side logic: hash-partition -> counts exchange -> all_to_all_single ->
local operator -> global row ids, and the two-phase group-by.  The union over ranks must equal the
single-table oracle result (multiset equality, the reference's own comparison rule)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


class OracleOps(object):
    """Same call surface as dist.GdfOps, computed by the CPU oracle on CPU tensors."""

    def hash_partition(self, cols, nparts):
        import oracle
        key = cols[0].numpy()
        pid = oracle.partition_ids([key], nparts)
        order = np.argsort(pid, kind="stable")
        counts = np.bincount(pid, minlength=nparts)
        offsets = np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(int).tolist()
        return [torch.from_numpy(c.numpy()[order].copy()) for c in cols], offsets

    def partition_pairs(self, keys, id_base, nparts):
        import oracle
        k = keys.numpy()
        pid = oracle.partition_ids([k], nparts)
        order = np.argsort(pid, kind="stable")
        counts = np.bincount(pid, minlength=nparts)
        offsets = np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(int).tolist()
        ids = (np.arange(len(k), dtype=np.int64) + id_base).astype(np.int32)
        return torch.from_numpy(k[order].copy()), torch.from_numpy(ids[order].copy()), offsets

    def hash_partition_rows(self, cols, num_keys, nparts):
        import oracle
        pid = oracle.partition_ids([c.numpy() for c in cols[:num_keys]], nparts)
        order = np.argsort(pid, kind="stable")
        counts = np.bincount(pid, minlength=nparts)
        offsets = np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(int).tolist()
        return [torch.from_numpy(c.numpy()[order].copy()) for c in cols], offsets

    def rows_valid_bytes(self, key_cols, valids):
        from oracle import np_oracle
        n = key_cols[0].numel()
        ok = np.ones(n, dtype=bool)
        for v in valids:
            ok &= np_oracle.unpack_valid(None if v is None else v.numpy(), n)
        return torch.from_numpy(ok.astype(np.int8))

    def left_join_masked(self, lkeys, lok, rkeys, rok, lids, rids):
        import oracle
        from oracle import np_oracle
        lv = [np_oracle.pack_valid(lok.numpy() != 0)] + [None] * (len(lkeys) - 1)
        rv = [np_oracle.pack_valid(rok.numpy() != 0)] + [None] * (len(rkeys) - 1)
        li, ri = oracle.join(oracle.JOIN_LEFT, [k.numpy() for k in lkeys], [k.numpy() for k in rkeys], lv, rv)
        lpn, rpn = lids.numpy(), rids.numpy()
        gl = np.where(li >= 0, lpn[np.maximum(li, 0)], -1).astype(np.int32) if len(li) else li
        gr = np.where(ri >= 0, rpn[np.maximum(ri, 0)], -1).astype(np.int32) if len(ri) else ri
        return torch.from_numpy(gl), torch.from_numpy(gr)

    def _join(self, kind, lk, rk, lp, rp):
        import oracle
        li, ri = oracle.join(kind, [lk.numpy()], [rk.numpy()])
        lpn, rpn = lp.numpy(), rp.numpy()
        gl = np.where(li >= 0, lpn[np.maximum(li, 0)], -1).astype(np.int32) if len(li) else li
        gr = np.where(ri >= 0, rpn[np.maximum(ri, 0)], -1).astype(np.int32) if len(ri) else ri
        return torch.from_numpy(gl), torch.from_numpy(gr)

    def inner_join(self, lk, rk, lp, rp):
        import oracle
        return self._join(oracle.JOIN_INNER, lk, rk, lp, rp)

    def left_join(self, lk, rk, lp, rp):
        import oracle
        return self._join(oracle.JOIN_LEFT, lk, rk, lp, rp)

    def group_by_sum(self, keys, vals):
        import oracle
        if keys.numel() == 0:
            return keys.clone(), vals.clone()
        k, a = oracle.groupby(oracle.OP_SUM, [keys.numpy()], vals.numpy())
        return torch.from_numpy(k[0]), torch.from_numpy(a)


def _c5_tables(rng):
    """BASELINE config C5 at test size: (int64,int32) key, 30 % null rows (one of the two masks cleared)."""
    nl, nr = 30_000, 3_000
    l = [rng.randint(0, nr, nl).astype(np.int64), rng.randint(0, 4, nl).astype(np.int32)]
    r = [rng.randint(0, nr, nr).astype(np.int64), rng.randint(0, 4, nr).astype(np.int32)]

    def masks(n):
        null_rows, which = rng.rand(n) < 0.3, rng.rand(n) < 0.5
        return [np.packbits(~(null_rows & which), bitorder="little"), np.packbits(~(null_rows & ~which), bitorder="little")]
    return l, r, masks(nl), masks(nr)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, case, tmpdir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from libgdf_b200 import dist as D
    ops = OracleOps()
    rng = np.random.RandomState(1234)          # same full tables on every rank, each takes its block
    try:
        if case in ("inner", "left"):
            P, B = 50_000, 7_001
            probe = rng.randint(0, 2 * B, P).astype(np.int64)
            build = rng.permutation(B).astype(np.int64)
            if case == "left":
                build = np.concatenate([build, build[:100]])       # some duplicate build keys
            plo, phi = D.shard_bounds(len(probe), world, rank)
            blo, bhi = D.shard_bounds(len(build), world, rank)
            gl, gr = D.distributed_join(case, torch.from_numpy(probe[plo:phi]), torch.from_numpy(build[blo:bhi]), plo, blo, ops)
            np.save(os.path.join(tmpdir, "l%d.npy" % rank), gl.numpy())
            np.save(os.path.join(tmpdir, "r%d.npy" % rank), gr.numpy())
        elif case == "c5":
            l, r, lv, rv = _c5_tables(rng)
            plo, phi = D.shard_bounds(len(l[0]), world, rank)
            blo, bhi = D.shard_bounds(len(r[0]), world, rank)

            def shard_mask(mask, lo, hi):   # re-pack the shard's bits (shards do not start on byte boundaries)
                from oracle import np_oracle
                return torch.from_numpy(np_oracle.pack_valid(np_oracle.unpack_valid(mask, len(mask) * 8)[lo:hi]))
            gl, gr = D.distributed_left_join_masked(
                [torch.from_numpy(c[plo:phi].copy()) for c in l], [shard_mask(m, plo, phi) for m in lv],
                [torch.from_numpy(c[blo:bhi].copy()) for c in r], [shard_mask(m, blo, bhi) for m in rv], plo, blo, ops)
            np.save(os.path.join(tmpdir, "l%d.npy" % rank), gl.numpy())
            np.save(os.path.join(tmpdir, "r%d.npy" % rank), gr.numpy())
        elif case == "groupby":
            N, G = 80_000, 500
            keys = (rng.zipf(1.3, N) % G).astype(np.int64) * 7919 + 13
            vals = rng.randint(0, 1000, N).astype(np.int64)
            lo, hi = D.shard_bounds(N, world, rank)
            gk, gv = D.distributed_group_by_sum(torch.from_numpy(keys[lo:hi]), torch.from_numpy(vals[lo:hi]), ops)
            np.save(os.path.join(tmpdir, "k%d.npy" % rank), gk.numpy())
            np.save(os.path.join(tmpdir, "v%d.npy" % rank), gv.numpy())
        elif case == "exchange":
            # ragged: rank r sends (p + 1) * (r + 1) rows to rank p, one partition empty
            counts = [(p + 1) * (rank + 1) if p != 1 or rank != 0 else 0 for p in range(world)]
            n = sum(counts)
            col = torch.arange(n, dtype=torch.int64) + 1000 * rank
            tag = torch.full((n,), rank, dtype=torch.int32)
            offsets = np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(int).tolist()
            (c, t), rc = D.exchange([col, tag], offsets)
            want_rc = [(rank + 1) * (s + 1) if rank != 1 or s != 0 else 0 for s in range(world)]
            assert rc == want_rc, (rc, want_rc)
            assert t.tolist() == sum([[s] * want_rc[s] for s in range(world)], [])
            assert c.numel() == sum(want_rc)
    finally:
        dist.destroy_process_group()


def _run(world, case, tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, case, str(tmp_path)), nprocs=world, join=True)


@pytest.mark.parametrize("world", [2, 3])
def test_exchange_ragged(world, tmp_path):
    _run(world, "exchange", tmp_path)


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("kind", ["inner", "left"])
def test_distributed_join_matches_single_table_oracle(world, kind, tmp_path):
    import oracle
    _run(world, kind, tmp_path)
    rng = np.random.RandomState(1234)
    P, B = 50_000, 7_001
    probe = rng.randint(0, 2 * B, P).astype(np.int64)
    build = rng.permutation(B).astype(np.int64)
    if kind == "left":
        build = np.concatenate([build, build[:100]])
    ol, orr = oracle.join(oracle.JOIN_INNER if kind == "inner" else oracle.JOIN_LEFT, [probe], [build])
    gl = np.concatenate([np.load(tmp_path / ("l%d.npy" % r)) for r in range(world)])
    gr = np.concatenate([np.load(tmp_path / ("r%d.npy" % r)) for r in range(world)])
    got = np.stack([gl, gr], 1)
    want = np.stack([ol, orr], 1)
    got = got[np.lexsort((got[:, 1], got[:, 0]))]
    want = want[np.lexsort((want[:, 1], want[:, 0]))]
    np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize("world", [2, 3])
def test_distributed_groupby_matches_single_table_oracle(world, tmp_path):
    import oracle
    _run(world, "groupby", tmp_path)
    rng = np.random.RandomState(1234)
    N, G = 80_000, 500
    keys = (rng.zipf(1.3, N) % G).astype(np.int64) * 7919 + 13
    vals = rng.randint(0, 1000, N).astype(np.int64)
    ok, oa = oracle.groupby(oracle.OP_SUM, [keys], vals)
    gk = np.concatenate([np.load(tmp_path / ("k%d.npy" % r)) for r in range(world)])
    gv = np.concatenate([np.load(tmp_path / ("v%d.npy" % r)) for r in range(world)])
    assert len(np.unique(gk)) == len(gk), "a key landed on two ranks"
    assert sorted(zip(gk.tolist(), gv.tolist())) == sorted(zip(ok[0].tolist(), oa.tolist()))


@pytest.mark.parametrize("world", [2, 3])
def test_distributed_c5_left_join_composite_key_with_nulls(world, tmp_path):
    import oracle
    _run(world, "c5", tmp_path)
    l, r, lv, rv = _c5_tables(np.random.RandomState(1234))
    ol, orr = oracle.join(oracle.JOIN_LEFT, l, r, lv, rv)
    gl = np.concatenate([np.load(tmp_path / ("l%d.npy" % k)) for k in range(world)])
    gr = np.concatenate([np.load(tmp_path / ("r%d.npy" % k)) for k in range(world)])
    got, want = np.stack([gl, gr], 1), np.stack([ol, orr], 1)
    np.testing.assert_array_equal(got[np.lexsort((got[:, 1], got[:, 0]))], want[np.lexsort((want[:, 1], want[:, 0]))])
    assert len(gl) >= len(l[0]) and (gr == -1).sum() > 0.25 * len(l[0])      # null / unmatched left rows are kept


# ---- the one-pass fused exchange: plan arithmetic (no GPU, no process group) ----
def test_fused_exchange_plan_is_a_consistent_layout():
    """Every rank derives its write offsets from the same gathered count matrix.  Simulate the scatter on numpy: the
    ranges of all senders must tile every receiver's buffer exactly, partition-major and without overlap, and the
    receiver's per-partition totals must describe that layout."""
    from libgdf_b200.dist import plan_fused_exchange
    rng = np.random.RandomState(5)
    for world, nlocal in ((2, 1), (2, 64), (3, 4), (8, 16)):
        bins = world * nlocal
        counts = rng.randint(0, 50, size=(world, bins)).tolist()
        counts[0][0] = 0                                                   # empty bins are legal
        plans = [plan_fused_exchange(counts, world, nlocal, r) for r in range(world)]
        recv_rows = plans[0][2]
        assert all(p[2] == recv_rows for p in plans)                       # same sizing on every rank
        for d in range(world):
            owner = -np.ones(recv_rows[d], dtype=np.int64)                 # which (partition, sender) wrote each slot
            for s in range(world):
                off = plans[s][0]
                for p in range(nlocal):
                    b = d * nlocal + p
                    sl = slice(off[b], off[b] + ((counts[s][b] + 3) & ~3))    # slots are padded to whole 32-byte granules (4 pairs)
                    assert (owner[sl] == -1).all(), "overlap"
                    owner[sl] = p * world + s
            assert (owner >= 0).all(), "hole"
            assert (np.diff(owner) >= 0).all()                             # partition-major, sender order inside
            totals = plans[d][1]
            assert totals == [int(((owner // world) == p).sum()) for p in range(nlocal)]


def test_fused_nlocal_choice():
    from libgdf_b200.dist import fused_nlocal
    assert fused_nlocal(100_000_000, 8) == 16        # C3 at 8 GPUs: 12.5 M build rows per rank
    assert fused_nlocal(100_000_000, 2) == 64
    assert fused_nlocal(100_000_000, 4) == 32
    assert fused_nlocal(1000, 8) == 1
    assert fused_nlocal(10 ** 10, 8) * 8 <= 256      # never more than 256 bins in one pass
