import sys, json, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from libgdf_b200.librmm_cffi import librmm, librmm_config
librmm_config.use_pool_allocator = True; librmm.finalize(); librmm.initialize()
from libgdf_b200.libgdf_cffi import ffi, libgdf_api as lib
from libgdf_b200 import columns as C
n_l, n_r = 100_000_000, 10_000_000
g = torch.Generator(device='cuda'); g.manual_seed(1)
l0 = torch.randint(0, n_r, (n_l,), generator=g, device='cuda', dtype=torch.int64); l1 = torch.randint(0, 4, (n_l,), generator=g, device='cuda', dtype=torch.int32)
r0 = torch.randint(0, n_r, (n_r,), generator=g, device='cuda', dtype=torch.int64); r1 = torch.randint(0, 4, (n_r,), generator=g, device='cuda', dtype=torch.int32)
ctx = ffi.new("gdf_context*"); lib.gdf_context_view(ctx, 0, lib.GDF_HASH, 0, 0, 0)
def mk(n):
    bits = (torch.rand(n, generator=g, device='cuda') < 0.7)
    pad = (-n) % 8
    b = torch.cat([bits, torch.zeros(pad, dtype=torch.bool, device='cuda')]).view(-1, 8).to(torch.uint8)
    w = torch.tensor([1, 2, 4, 8, 16, 32, 64, 128], dtype=torch.uint8, device='cuda')
    return (b * w).sum(1).to(torch.uint8)
import os
if os.environ.get('MASK'):
    L = [C.Column(l0, mk(n_l)), C.Column(l1)]; R = [C.Column(r0, mk(n_r)), C.Column(r1)]
else:
    L = [C.Column(l0), C.Column(l1)]; R = [C.Column(r0), C.Column(r1)]
idx = ffi.new("int[]", [0, 1])
lib.gdfx_profile_enable(1)
for it in range(3):
    ol, orr = ffi.new("gdf_column*"), ffi.new("gdf_column*")
    torch.cuda.synchronize(); t0 = time.perf_counter()
    rc = lib.gdf_left_join(C.column_array(L), 2, idx, C.column_array(R), 2, idx, 2, 0, ffi.NULL, ol, orr, ctx)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print('rc', rc, 'pairs', int(ol.size), 'ms %.2f' % ((t1 - t0) * 1e3))
    lib.gdf_column_free(ol); lib.gdf_column_free(orr)
buf = ffi.new("char[]", 1 << 16); lib.gdfx_profile_report(buf, 1 << 16); print(ffi.string(buf).decode())
