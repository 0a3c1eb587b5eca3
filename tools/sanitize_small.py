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
