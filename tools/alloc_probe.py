"""How long do stream-ordered pool allocations of join-sized blocks take on the host? (diagnostic)"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.cuda.init()
from libgdf_b200.librmm_cffi import librmm, librmm_config, ffi, librmm_api
librmm_config.use_pool_allocator = True
librmm.finalize(); librmm.initialize()
sizes = [int(8e9), int(4e9), int(4e9), int(8e8), int(4e9), int(4e9)]
for rep in range(4):
    ptrs = []
    t0 = time.perf_counter()
    for s in sizes:
        p = ffi.new("void**")
        librmm_api.rmmAlloc(p, s, ffi.NULL)
        ptrs.append(p[0])
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    for p in ptrs:
        librmm_api.rmmFree(p, ffi.NULL)
    t2 = time.perf_counter()
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    print("rep %d: alloc %.2f ms  free %.2f ms  sync %.2f ms" % (rep, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3))
# varying order / interleaved syncs like the join does
for rep in range(3):
    t0 = time.perf_counter()
    a = ffi.new("void**"); librmm_api.rmmAlloc(a, int(8e9), ffi.NULL)
    b = ffi.new("void**"); librmm_api.rmmAlloc(b, int(4e9), ffi.NULL)
    torch.cuda.synchronize()
    c = ffi.new("void**"); librmm_api.rmmAlloc(c, int(4.2e9), ffi.NULL)
    librmm_api.rmmFree(a[0], ffi.NULL)
    torch.cuda.synchronize()
    d = ffi.new("void**"); librmm_api.rmmAlloc(d, int(4e9), ffi.NULL)
    e = ffi.new("void**"); librmm_api.rmmAlloc(e, int(4e9), ffi.NULL)
    for p in (b, c, d, e):
        librmm_api.rmmFree(p[0], ffi.NULL)
    torch.cuda.synchronize()
    print("interleaved rep %d: %.2f ms" % (rep, (time.perf_counter() - t0) * 1e3))
