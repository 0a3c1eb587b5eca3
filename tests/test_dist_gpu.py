"""-m gpu: the per-shard operators of the multi-GPU layer (libgdf_b200.dist.GdfOps: gdf_hash_partition,
gdf_inner_join / gdf_left_join + gdfx_remap_indices, gdf_group_by_sum) on ONE GPU, with the all-to-all
emulated in-process: R virtual ranks partition their shards, the partitions are regrouped by
destination exactly as dist.exchange would deliver them, and every virtual rank runs its local operator.
The union must equal the single-table oracle result.  (The real exchange is covered by
tests/test_dist_cpu.py under gloo, and by bench.py --gpus N under NCCL.)"""
import numpy as np
import pytest
import torch

import oracle
from libgdf_b200 import dist as D

pytestmark = pytest.mark.gpu


def _emulated_exchange(per_rank_cols, per_rank_offsets, world):
    """per_rank_cols[r] = partition-contiguous columns of rank r -> received columns per destination."""
    recv = []
    for dst in range(world):
        cols = []
        for c in range(len(per_rank_cols[0])):
            parts = []
            for src in range(world):
                n = per_rank_cols[src][c].numel()
                b = list(per_rank_offsets[src]) + [n]
                parts.append(per_rank_cols[src][c][b[dst]:b[dst + 1]])
            cols.append(torch.cat(parts))
        recv.append(cols)
    return recv


@pytest.mark.parametrize("world", [2, 8])
@pytest.mark.parametrize("kind", ["inner", "left"])
def test_sharded_join_matches_oracle(world, kind):
    ops = D.GdfOps()
    P, B = 300_000, 50_000
    probe = np.random.randint(0, 2 * B, P).astype(np.int64)
    build = np.random.permutation(B).astype(np.int64)
    if kind == "left":
        build = np.concatenate([build, build[:500]])
    shards_l, shards_r, offs_l, offs_r = [], [], [], []
    for r in range(world):
        plo, phi = D.shard_bounds(len(probe), world, r)
        blo, bhi = D.shard_bounds(len(build), world, r)
        lk = torch.from_numpy(probe[plo:phi]).cuda()
        rk = torch.from_numpy(build[blo:bhi]).cuda()
        k1, i1, ol = ops.partition_pairs(lk, plo, world)
        k2, i2, orr = ops.partition_pairs(rk, blo, world)
        # every row keeps its key/id pairing and lands in exactly one destination range
        assert sorted(zip(k1.cpu().tolist(), i1.cpu().tolist())) == sorted(zip(probe[plo:phi].tolist(), range(plo, phi)))
        shards_l.append([k1, i1]), offs_l.append(ol), shards_r.append([k2, i2]), offs_r.append(orr)
    recv_l = _emulated_exchange(shards_l, offs_l, world)
    recv_r = _emulated_exchange(shards_r, offs_r, world)
    got_l, got_r = [], []
    for r in range(world):
        fn = ops.inner_join if kind == "inner" else ops.left_join
        gl, gr = fn(recv_l[r][0], recv_r[r][0], recv_l[r][1], recv_r[r][1])
        got_l.append(gl.cpu().numpy()), got_r.append(gr.cpu().numpy())
    gl, gr = np.concatenate(got_l), np.concatenate(got_r)
    ol, orr = oracle.join(oracle.JOIN_INNER if kind == "inner" else oracle.JOIN_LEFT, [probe], [build])
    got = np.stack([gl, gr], 1)
    want = np.stack([ol, orr], 1)
    np.testing.assert_array_equal(got[np.lexsort((got[:, 1], got[:, 0]))], want[np.lexsort((want[:, 1], want[:, 0]))])


@pytest.mark.parametrize("world", [2, 8])
def test_sharded_groupby_matches_oracle(world):
    ops = D.GdfOps()
    N, G = 600_000, 3000
    keys = (np.random.zipf(1.2, N) % G).astype(np.int64) * 7919 + 13
    vals = np.random.randint(0, 1000, N).astype(np.int64)
    partial_cols, partial_offs = [], []
    for r in range(world):
        lo, hi = D.shard_bounds(N, world, r)
        pk, pv = ops.group_by_sum(torch.from_numpy(keys[lo:hi]).cuda(), torch.from_numpy(vals[lo:hi]).cuda())
        cols, off = ops.hash_partition([pk.contiguous(), pv.contiguous()], world)
        partial_cols.append(cols), partial_offs.append(off)
    recv = _emulated_exchange(partial_cols, partial_offs, world)
    gk, gv = [], []
    for r in range(world):
        k, v = ops.group_by_sum(recv[r][0], recv[r][1])
        gk.append(k.cpu().numpy()), gv.append(v.cpu().numpy())
    gk, gv = np.concatenate(gk), np.concatenate(gv)
    ok, oa = oracle.groupby(oracle.OP_SUM, [keys], vals)
    assert len(np.unique(gk)) == len(gk), "a key landed on two ranks"
    assert sorted(zip(gk.tolist(), gv.tolist())) == sorted(zip(ok[0].tolist(), oa.tolist()))


def test_remap_indices_keeps_minus_one():
    from libgdf_b200 import columns as C
    from libgdf_b200.libgdf_cffi import ffi, libgdf
    idx = torch.tensor([0, -1, 3, 2, -1, 1], dtype=torch.int32, device="cuda")
    payload = torch.tensor([100, 101, 102, 103], dtype=torch.int32, device="cuda")
    col = C.Column(idx)
    libgdf.gdfx_remap_indices(col.cdata, ffi.cast("int32_t*", payload.data_ptr()), payload.numel())
    torch.cuda.synchronize()
    assert idx.cpu().tolist() == [100, -1, 103, 102, -1, 101]


@pytest.mark.parametrize("world", [2, 5, 8])
def test_partition_scatter_peer_matches_partition_pairs(world):
    """The fused partition+exchange kernel with LOCAL destination buffers (the peer pointers of a real run
    are ordinary device addresses to the kernel): every destination receives exactly the range that
    gdfx_partition_pairs assigns to it, at the given offset."""
    ops = D.GdfOps()
    n, base = 700_003, 1000
    keys = torch.from_numpy(np.random.randint(0, 1 << 40, n).astype(np.int64)).cuda()
    pk, pi, offs = ops.partition_pairs(keys, base, world)
    counts = ops.partition_count(keys, world)
    bounds = offs + [n]
    assert counts == [bounds[p + 1] - bounds[p] for p in range(world)]
    pad = 17                                                     # this "rank" writes behind 17 foreign rows
    dk = [torch.full((counts[p] + pad,), -7, dtype=torch.int64, device="cuda") for p in range(world)]
    di = [torch.full((counts[p] + pad,), -7, dtype=torch.int32, device="cuda") for p in range(world)]
    ops.partition_scatter_peer(keys, base, [t.data_ptr() for t in dk], [t.data_ptr() for t in di], [pad] * world)
    torch.cuda.synchronize()
    for p in range(world):
        assert (dk[p][:pad] == -7).all() and (di[p][:pad] == -7).all()
        got = sorted(zip(dk[p][pad:].cpu().tolist(), di[p][pad:].cpu().tolist()))
        want = sorted(zip(pk[bounds[p]:bounds[p + 1]].cpu().tolist(), pi[bounds[p]:bounds[p + 1]].cpu().tolist()))
        assert got == want


@pytest.mark.parametrize("world", [2, 8])
def test_sharded_c5_left_join_composite_key_with_nulls(world):
    """BASELINE config C5 at test size through the multi-GPU operators (row validity travels as a byte
    column, gdf_hash_partition on both key columns, gdf_left_join on the rebuilt mask), exchange emulated."""
    from oracle import np_oracle
    ops = D.GdfOps()
    nl, nr = 200_000, 20_000
    l = [np.random.randint(0, nr, nl).astype(np.int64), np.random.randint(0, 4, nl).astype(np.int32)]
    r = [np.random.randint(0, nr, nr).astype(np.int64), np.random.randint(0, 4, nr).astype(np.int32)]

    def masks(n):
        null_rows, which = np.random.rand(n) < 0.3, np.random.rand(n) < 0.5
        return [np.packbits(~(null_rows & which), bitorder="little"), np.packbits(~(null_rows & ~which), bitorder="little")]
    lv, rv = masks(nl), masks(nr)

    def shard(cols, valids, lo, hi):
        keys = [torch.from_numpy(c[lo:hi].copy()).cuda() for c in cols]
        vs = [torch.from_numpy(np_oracle.pack_valid(np_oracle.unpack_valid(m, len(m) * 8)[lo:hi])).cuda() for m in valids]
        ok = ops.rows_valid_bytes(keys, vs)
        want_ok = np.ones(hi - lo, dtype=bool)
        for m in valids:
            want_ok &= np_oracle.unpack_valid(m, len(m) * 8)[lo:hi]
        np.testing.assert_array_equal(ok.cpu().numpy().astype(bool), want_ok)
        ids = torch.arange(lo, hi, dtype=torch.int32, device="cuda")
        return ops.hash_partition_rows(keys + [ids, ok], 2, world)

    lparts, rparts = [], []
    for k in range(world):
        lparts.append(shard(l, lv, *D.shard_bounds(nl, world, k)))
        rparts.append(shard(r, rv, *D.shard_bounds(nr, world, k)))
    recv_l = _emulated_exchange([p[0] for p in lparts], [p[1] for p in lparts], world)
    recv_r = _emulated_exchange([p[0] for p in rparts], [p[1] for p in rparts], world)
    gl, gr = [], []
    for k in range(world):
        a, b = ops.left_join_masked(recv_l[k][:2], recv_l[k][3], recv_r[k][:2], recv_r[k][3], recv_l[k][2], recv_r[k][2])
        gl.append(a.cpu().numpy()), gr.append(b.cpu().numpy())
    gl, gr = np.concatenate(gl), np.concatenate(gr)
    ol, orr = oracle.join(oracle.JOIN_LEFT, l, r, lv, rv)
    got, want = np.stack([gl, gr], 1), np.stack([ol, orr], 1)
    np.testing.assert_array_equal(got[np.lexsort((got[:, 1], got[:, 0]))], want[np.lexsort((want[:, 1], want[:, 0]))])


@pytest.mark.parametrize("world,nlocal", [(2, 1), (2, 8), (3, 4), (8, 16), (8, 32)])
@pytest.mark.parametrize("key_dtype", [np.int64, np.int32])
def test_one_pass_fused_exchange_matches_oracle(world, nlocal, key_dtype):
    """PeerExchange.fused_inner_join on ONE GPU: `world` virtual ranks count with the combined (destination rank x
    receiver-local partition) geometry, derive the plan from the gathered count matrix, scatter compact pairs
    straight into the destinations' buffers (local tensors here; IPC-mapped peer memory in a real run) and every
    virtual rank joins what it received WITHOUT partitioning again.  Union == the single-table oracle join."""
    ops = D.GdfOps()
    P, B = 400_003, 60_000
    probe = np.random.randint(0, 2 * B, P).astype(key_dtype)
    build = np.random.permutation(B).astype(key_dtype)
    bins = world * nlocal
    shards, counts_b, counts_p = [], [], []
    for r in range(world):
        plo, phi = D.shard_bounds(P, world, r)
        blo, bhi = D.shard_bounds(B, world, r)
        pk, bk = torch.from_numpy(probe[plo:phi]).cuda(), torch.from_numpy(build[blo:bhi]).cuda()
        cb, hi_b = ops.xjoin_count(bk, world, nlocal)
        cp, _ = ops.xjoin_count(pk, world, nlocal)
        assert hi_b == 0 and sum(cb) == bhi - blo and sum(cp) == phi - plo and len(cb) == bins
        shards.append((pk, plo, bk, blo)), counts_b.append(cb), counts_p.append(cp)
    plans_b = [D.plan_fused_exchange(counts_b, world, nlocal, r) for r in range(world)]
    plans_p = [D.plan_fused_exchange(counts_p, world, nlocal, r) for r in range(world)]
    pad = 5                                                      # sentinel pairs behind every receive buffer
    buf_b = [torch.full((plans_b[0][2][d] + pad, 2), -7, dtype=torch.int32, device="cuda") for d in range(world)]
    buf_p = [torch.full((plans_p[0][2][d] + pad, 2), -7, dtype=torch.int32, device="cuda") for d in range(world)]
    for r, (pk, plo, bk, blo) in enumerate(shards):
        ops.xjoin_scatter(bk, blo, world, nlocal, [t.data_ptr() for t in buf_b], plans_b[r][0], counts_b[r])
        ops.xjoin_scatter(pk, plo, world, nlocal, [t.data_ptr() for t in buf_p], plans_p[r][0], counts_p[r])
    torch.cuda.synchronize()
    got_l, got_r = [], []
    for d in range(world):
        assert (buf_b[d][-pad:] == -7).all() and (buf_p[d][-pad:] == -7).all()          # nothing written past the plan
        INT_MIN = -(1 << 31)                                                             # the "no row" pad of an odd slot
        for buf in (buf_b[d], buf_p[d]):                                                  # no hole left inside the plan
            tags = buf[:-pad, 1]
            assert bool(((tags >= 0) | (tags == INT_MIN)).all()) and int((tags == INT_MIN).sum()) <= 3 * world * nlocal
        gl, gr = ops.xjoin_local(buf_p[d].data_ptr(), plans_p[d][1], buf_b[d].data_ptr(), plans_b[d][1], nlocal)
        got_l.append(gl.cpu().numpy()), got_r.append(gr.cpu().numpy())
    gl, gr = np.concatenate(got_l), np.concatenate(got_r)
    ol, orr = oracle.join(oracle.JOIN_INNER, [probe.astype(np.int64)], [build.astype(np.int64)])
    got, want = np.stack([gl, gr], 1), np.stack([ol, orr], 1)
    np.testing.assert_array_equal(got[np.lexsort((got[:, 1], got[:, 0]))], want[np.lexsort((want[:, 1], want[:, 0]))])


def test_one_pass_fused_exchange_reports_wide_keys():
    """A build key that does not fit 32 bits makes the compact exchange inapplicable: xjoin_count says so (hi_or != 0)
    and PeerExchange falls back to the two-pass path."""
    ops = D.GdfOps()
    keys = torch.tensor([1, 2, (1 << 40) + 3, 4], dtype=torch.int64, device="cuda")
    counts, hi = ops.xjoin_count(keys, 2, 4)
    assert hi != 0 and sum(counts) == 3          # the wide key is counted nowhere, exactly as the scatter drops it


@pytest.mark.parametrize("world,nlocal", [(2, 8), (3, 4), (8, 32)])
def test_device_side_exchange_plan_equals_host_plan(world, nlocal):
    """The asynchronous route of the one-pass exchange keeps counts and plan on the device (gdfx_xjoin_count_dev /
    _plan_dev / _scatter_dev).  Same counts, same offsets as the host plan, same bytes in the receive buffers; and the
    status flags stop the scatter (nothing written) when a buffer is too small or a build key is wide."""
    ops = D.GdfOps()
    P, B = 300_007, 40_000
    probe = np.random.randint(0, 2 * B, P).astype(np.int64)
    build = np.random.permutation(B).astype(np.int64)
    bins = world * nlocal
    stride = 2 * (bins + 1)
    shards, allc = [], torch.zeros(world * stride, dtype=torch.int64, device="cuda")
    for r in range(world):
        plo, phi = D.shard_bounds(P, world, r)
        blo, bhi = D.shard_bounds(B, world, r)
        pk, bk = torch.from_numpy(probe[plo:phi]).cuda(), torch.from_numpy(build[blo:bhi]).cuda()
        mine = allc[r * stride:(r + 1) * stride]
        ops.xjoin_count_dev(bk, world, nlocal, mine[:bins + 1])
        ops.xjoin_count_dev(pk, world, nlocal, mine[bins + 1:])
        cb, hi_b = ops.xjoin_count(bk, world, nlocal)
        cp, _ = ops.xjoin_count(pk, world, nlocal)
        assert mine.cpu().tolist() == cb + [hi_b] + cp + [0]
        shards.append((pk, plo, bk, blo))
    M = allc.view(world, stride).cpu().numpy()
    counts_b, counts_p = M[:, :bins].tolist(), M[:, bins + 1:2 * bins + 1].tolist()
    recv_b = D.plan_fused_exchange(counts_b, world, nlocal, 0)[2]
    recv_p = D.plan_fused_exchange(counts_p, world, nlocal, 0)[2]
    cap_b, cap_p = max(recv_b), max(recv_p)
    buf_b = [torch.full((cap_b, 2), -7, dtype=torch.int32, device="cuda") for _ in range(world)]
    buf_p = [torch.full((cap_p, 2), -7, dtype=torch.int32, device="cuda") for _ in range(world)]
    off = torch.zeros(2 * bins, dtype=torch.int64, device="cuda")
    status = torch.zeros(2, dtype=torch.int32, device="cuda")
    for r, (pk, plo, bk, blo) in enumerate(shards):
        # too small a buffer: flagged, and the scatter leaves the buffers untouched
        ops.xjoin_plan_dev(allc, world, nlocal, r, cap_b, cap_p - 1, off[:bins], off[bins:], status)
        assert status.cpu().tolist() == [0, 1]
        mine_r = allc[r * stride:(r + 1) * stride]
        if r == 0:                                   # nothing has been written yet
            ops.xjoin_scatter_dev(pk, plo, world, nlocal, [t.data_ptr() for t in buf_p], off[bins:], mine_r[bins + 1:2 * bins + 1], status)
            assert all(bool((t == -7).all()) for t in buf_p)
        ops.xjoin_plan_dev(allc, world, nlocal, r, cap_b, cap_p, off[:bins], off[bins:], status)
        assert status.cpu().tolist() == [0, 0]
        want_b, want_p = D.plan_fused_exchange(counts_b, world, nlocal, r)[0], D.plan_fused_exchange(counts_p, world, nlocal, r)[0]
        assert off.cpu().tolist() == want_b + want_p
        ops.xjoin_scatter_dev(bk, blo, world, nlocal, [t.data_ptr() for t in buf_b], off[:bins], mine_r[:bins], status)
        ops.xjoin_scatter_dev(pk, plo, world, nlocal, [t.data_ptr() for t in buf_p], off[bins:], mine_r[bins + 1:2 * bins + 1], status)
    torch.cuda.synchronize()
    got_l, got_r = [], []
    for d in range(world):
        tot_b = D.plan_fused_exchange(counts_b, world, nlocal, d)[1]
        tot_p = D.plan_fused_exchange(counts_p, world, nlocal, d)[1]
        for buf, rows in ((buf_b[d], recv_b[d]), (buf_p[d], recv_p[d])):              # filled completely: real pairs or pad pairs
            tags = buf[:rows, 1]
            assert bool(((tags >= 0) | (tags == -(1 << 31))).all())
        if d % 2:      # the two-stage form: tables filled on the library's private stream, then the probe waits for them
            handle = ops.xjoin_build(buf_b[d].data_ptr(), tot_b, nlocal, True)
            gl, gr = ops.xjoin_probe(handle, buf_p[d].data_ptr(), tot_p)
        else:
            gl, gr = ops.xjoin_local(buf_p[d].data_ptr(), tot_p, buf_b[d].data_ptr(), tot_b, nlocal)
        got_l.append(gl.cpu().numpy()), got_r.append(gr.cpu().numpy())
    gl, gr = np.concatenate(got_l), np.concatenate(got_r)
    ol, orr = oracle.join(oracle.JOIN_INNER, [probe], [build])
    got, want = np.stack([gl, gr], 1), np.stack([ol, orr], 1)
    np.testing.assert_array_equal(got[np.lexsort((got[:, 1], got[:, 0]))], want[np.lexsort((want[:, 1], want[:, 0]))])
    # a wide build key anywhere: flagged
    allc[bins] = 5
    ops.xjoin_plan_dev(allc, world, nlocal, 0, cap_b, cap_p, off[:bins], off[bins:], status)
    assert status.cpu().tolist()[0] == 1
