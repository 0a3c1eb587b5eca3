"""Small end-to-end pass over the default kernels, meant to be run under compute-sanitizer:
   compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch
import oracle
import gpu_utils as G
from libgdf_b200 import columns as C
from libgdf_b200.libgdf_cffi import ffi, libgdf

rng = np.random.RandomState(7)
# gdf_filter: chunked streaming select, ragged tail
col = rng.randint(0, 10, 300_017).astype(np.int64)
c = C.column(col)
d_cols = torch.zeros(1, dtype=torch.int64, device="cuda"); d_types = torch.zeros(1, dtype=torch.int32, device="cuda")
val = torch.tensor([3], dtype=torch.int64, device="cuda"); d_vals = torch.tensor([val.data_ptr()], dtype=torch.int64, device="cuda")
d_indx = torch.zeros(len(col), dtype=torch.int64, device="cuda"); new_sz = ffi.new("size_t*")
libgdf.gdf_filter(len(col), C.struct_array([c]), 1, ffi.cast("void**", d_cols.data_ptr()), ffi.cast("int*", d_types.data_ptr()),
                  ffi.cast("void**", d_vals.data_ptr()), ffi.cast("size_t*", d_indx.data_ptr()), new_sz)
assert np.array_equal(d_indx[: int(new_sz[0])].cpu().numpy(), np.nonzero(col == 3)[0]); print("filter ok")
# group-by fast path
keys = rng.randint(0, 5000, 300_011).astype(np.int64); vals = rng.randint(0, 1000, 300_011).astype(np.int64)
gk, ga = G.groupby("sum", [keys], vals); ok, oa = oracle.groupby(oracle.OP_SUM, [keys], vals)
assert G.rows_as_sorted_tuples(gk, ga) == G.rows_as_sorted_tuples(ok, oa); print("groupby ok")
# partitioned join, single and composite key
b = rng.permutation(1_100_000).astype(np.int64); p = rng.randint(0, 1_300_000, 1_500_003).astype(np.int64)
gl, gr = G.join("inner", [p], [b]); ol, orr = oracle.join(oracle.JOIN_INNER, [p], [b])
assert np.array_equal(G.sorted_pairs(gl, gr), G.sorted_pairs(ol, orr)); print("join ok")
b2 = [rng.randint(0, 600_000, 1_100_000).astype(np.int64), rng.randint(0, 4, 1_100_000).astype(np.int32)]
p2 = [rng.randint(0, 600_000, 1_200_001).astype(np.int64), rng.randint(0, 4, 1_200_001).astype(np.int32)]
gl, gr = G.join("left", p2, b2); ol, orr = oracle.join(oracle.JOIN_LEFT, p2, b2)
assert np.array_equal(G.sorted_pairs(gl, gr), G.sorted_pairs(ol, orr)); print("composite join ok")
# ---- round 2 kernels ----
# LEFT join on the compact path (element-parallel scatter with NULL tags, positional probe) and a skewed INNER join
# (region overflow -> exact fallback of the histogram-free probe side)
gl, gr = G.join("left", [p], [b]); ol, orr = oracle.join(oracle.JOIN_LEFT, [p], [b])
assert np.array_equal(G.sorted_pairs(gl, gr), G.sorted_pairs(ol, orr)); print("compact left join ok")
ps = p.copy(); ps[rng.rand(len(ps)) < 0.5] = 4242
gl, gr = G.join("inner", [ps], [b]); ol, orr = oracle.join(oracle.JOIN_INNER, [ps], [b])
assert np.array_equal(G.sorted_pairs(gl, gr), G.sorted_pairs(ol, orr)); print("skewed inner join (overflow fallback) ok")
# gpu_apply_stencil through the chunked select (dense + sparse emission) and multi-column gdf_filter
st_col = rng.randint(0, 10, 200_003).astype(np.int64)
lhs, stc, outc = C.column(st_col), C.empty_column(len(st_col), torch.int8, with_valid=True), C.empty_column(len(st_col), torch.int64, with_valid=True)
libgdf.gpu_comparison_static_i64(lhs.cdata, 3, stc.cdata, libgdf.GDF_EQUALS)
libgdf.gpu_apply_stencil(lhs.cdata, stc.cdata, outc.cdata)
assert int(outc.cdata.size) == int((st_col == 3).sum()); print("stencil ok")
# result_cols gather, hash partition
L = [C.column(np.arange(len(p), dtype=np.int32)), C.column(p)]; R = [C.column(b), C.column(np.arange(len(b), dtype=np.float64))]
res = [ffi.new("gdf_column*") for _ in range(3)]; ctx = ffi.new("gdf_context*"); libgdf.gdf_context_view(ctx, 0, libgdf.GDF_HASH, 0, 0, 0)
o_l, o_r = ffi.new("gdf_column*"), ffi.new("gdf_column*")
libgdf.gdf_left_join(C.column_array(L), 2, ffi.new("int[]", [1]), C.column_array(R), 2, ffi.new("int[]", [0]), 1, 3, ffi.new("gdf_column*[]", res), o_l, o_r, ctx)
torch.cuda.synchronize()
for cc in res + [o_l, o_r]:
    libgdf.gdf_column_free(cc)
print("result_cols ok")
hk = C.column(rng.randint(0, 1 << 40, 300_001).astype(np.int64)); hv = C.column(np.arange(300_001, dtype=np.int64))
ok_, ov_ = C.empty_column(300_001, torch.int64), C.empty_column(300_001, torch.int64)
offs = ffi.new("int[]", 8)
libgdf.gdf_hash_partition(2, C.column_array([hk, hv]), ffi.new("int[]", [0]), 1, 8, C.column_array([ok_, ov_]), offs, libgdf.GDF_HASH_MURMUR3)
assert int(ov_.data.sum().item()) == 300_001 * 300_000 // 2; print("hash partition ok")
# the multi-GPU exchange on one GPU: 3 virtual ranks, sector-aligned scatter into local "peer" buffers, two-stage local join
from libgdf_b200 import dist as D
ops = D.GdfOps(); world, nlocal = 3, 4; bins = world * nlocal
P_, B_ = 120_003, 20_000
pr_ = rng.randint(0, 2 * B_, P_).astype(np.int64); bd_ = rng.permutation(B_).astype(np.int64)
sh, cb_, cp_ = [], [], []
for r in range(world):
    plo, phi = D.shard_bounds(P_, world, r); blo, bhi = D.shard_bounds(B_, world, r)
    pk, bk = torch.from_numpy(pr_[plo:phi]).cuda(), torch.from_numpy(bd_[blo:bhi]).cuda()
    cb_.append(ops.xjoin_count(bk, world, nlocal)[0]); cp_.append(ops.xjoin_count(pk, world, nlocal)[0]); sh.append((pk, plo, bk, blo))
pb = [D.plan_fused_exchange(cb_, world, nlocal, r) for r in range(world)]; pp_ = [D.plan_fused_exchange(cp_, world, nlocal, r) for r in range(world)]
bb = [torch.zeros((pb[0][2][d], 2), dtype=torch.int32, device="cuda") for d in range(world)]
bp_ = [torch.zeros((pp_[0][2][d], 2), dtype=torch.int32, device="cuda") for d in range(world)]
for r, (pk, plo, bk, blo) in enumerate(sh):
    ops.xjoin_scatter(bk, blo, world, nlocal, [t.data_ptr() for t in bb], pb[r][0], cb_[r])
    ops.xjoin_scatter(pk, plo, world, nlocal, [t.data_ptr() for t in bp_], pp_[r][0], cp_[r])
got = []
for d in range(world):
    h = ops.xjoin_build(bb[d].data_ptr(), pb[d][1], nlocal, True)
    a, c_ = ops.xjoin_probe(h, bp_[d].data_ptr(), pp_[d][1]); got.append((a.cpu().numpy(), c_.cpu().numpy()))
gl = np.concatenate([g[0] for g in got]); gr = np.concatenate([g[1] for g in got])
ol, orr = oracle.join(oracle.JOIN_INNER, [pr_], [bd_])
assert np.array_equal(G.sorted_pairs(gl, gr), G.sorted_pairs(ol, orr)); print("one-pass exchange ok")
