"""-m gpu: behaviour of librmm.so, both allocation modes.  Port of the reference's python/tests/test_rmm.py
(alloc / copy round trip over the same element counts, CSV log header) plus the C-level entry points the reference
declares in include/memory.h:65-184: rmmRealloc, rmmGetInfo, rmmGetAllocationOffset, rmmWriteLog / rmmLogSize /
rmmGetLog, stream arguments, and the extensions libgdf.so relies on (pool trimming)."""
import os

import numpy as np
import pytest
import torch

from gpu_utils import gen_rand

pytestmark = pytest.mark.gpu
HEADER = "Event Type,Device ID,Address,Stream,Size (bytes),Free Memory,Total Memory,Current Allocs,Start,End,Elapsed"


@pytest.fixture(params=["pool", "default"])
def rmm(request):
    from libgdf_b200.librmm_cffi import librmm, librmm_config
    old = (librmm_config.use_pool_allocator, librmm_config.enable_logging)
    librmm.finalize()
    librmm_config.use_pool_allocator = request.param == "pool"
    librmm_config.enable_logging = True
    librmm.initialize()
    yield librmm
    librmm.finalize()
    librmm_config.use_pool_allocator, librmm_config.enable_logging = old
    librmm.initialize()


@pytest.mark.parametrize("nelem", [1, 2, 7, 8, 9, 32, 128])
def test_rmm_alloc(rmm, nelem):                                # reference test_rmm.py:16-33
    h_in = gen_rand(np.int32, nelem)
    d_in = rmm.to_device(h_in)
    d_result = rmm.device_array_like(d_in)
    d_result.copy_(d_in)
    np.testing.assert_array_equal(d_result.cpu().numpy(), h_in)


def test_rmm_csv_log(rmm):                                     # reference test_rmm.py:35-54
    h_in = gen_rand(np.int32, 1024)
    d_in = rmm.to_device(h_in)
    d_result = rmm.device_array_like(d_in)
    d_result.copy_(d_in)
    del d_in, d_result
    csv = rmm.csv_log()
    assert csv.find(HEADER) >= 0
    lines = csv.strip().splitlines()
    kinds = [l.split(",")[0] for l in lines[1:]]
    assert kinds.count("Alloc") >= 2 and kinds.count("Free") >= 2
    sizes = [int(l.split(",")[4]) for l in lines[1:] if l.startswith("Alloc")]
    assert 4096 in sizes


def _ptr(rmm):
    return rmm._ffi.new("void **")


def test_alloc_free_realloc_and_streams(rmm):
    ffi = rmm._ffi
    s = torch.cuda.Stream()
    stream = ffi.cast("cudaStream_t", s.cuda_stream)
    p = _ptr(rmm)
    rmm.rmmAlloc(p, 1 << 20, stream)
    first = int(ffi.cast("uintptr_t", p[0]))
    assert first and first % 256 == 0
    with torch.cuda.stream(s):
        a = torch.as_tensor(_Alias(first, 1 << 18), device="cuda")
        a.fill_(7)
    rmm.rmmRealloc(p, 4 << 20, stream)                         # contents are NOT preserved (memory.cpp:172-194)
    second = int(ffi.cast("uintptr_t", p[0]))
    assert second and second % 256 == 0
    with torch.cuda.stream(s):
        b = torch.as_tensor(_Alias(second, 1 << 20), device="cuda")
        b.fill_(9)
        assert int(b.sum().item()) == 9 * (1 << 20)
    rmm.rmmFree(p[0], stream)
    rmm.rmmFree(ffi.NULL, stream)                              # freeing NULL is a no-op
    rmm.rmmAlloc(ffi.NULL, 0, stream)                          # (nullptr, 0) succeeds in the reference too
    with pytest.raises(Exception) as e:
        rmm.rmmAlloc(ffi.NULL, 16, stream)
    assert "RMM_ERROR_INVALID_ARGUMENT" in str(e.value)


def test_block_freed_on_one_stream_is_safe_to_reuse_on_another(rmm):
    """the pool must not hand a block freed on stream A to stream B while A still uses it."""
    ffi = rmm._ffi
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
    n = 1 << 24
    p = _ptr(rmm)
    rmm.rmmAlloc(p, n * 4, ffi.cast("cudaStream_t", sa.cuda_stream))
    addr = int(ffi.cast("uintptr_t", p[0]))
    with torch.cuda.stream(sa):
        a = torch.as_tensor(_Alias(addr, n), device="cuda")
        a.zero_()
        for _ in range(20):                                    # keep stream A busy on the block
            a.add_(1)
        checksum = a.sum()
    rmm.rmmFree(p[0], ffi.cast("cudaStream_t", sa.cuda_stream))
    q = _ptr(rmm)
    rmm.rmmAlloc(q, n * 4, ffi.cast("cudaStream_t", sb.cuda_stream))
    with torch.cuda.stream(sb):
        b = torch.as_tensor(_Alias(int(ffi.cast("uintptr_t", q[0])), n), device="cuda")
        b.fill_(-1)                                            # would corrupt A's sum if it ran early on a reused block
    torch.cuda.synchronize()
    assert int(checksum.item()) == 20 * n                      # B's fill did not overtake A's work on a reused block
    rmm.rmmFree(q[0], ffi.cast("cudaStream_t", sb.cuda_stream))


def test_get_info_and_allocation_offset(rmm):
    ffi = rmm._ffi
    free_b, total_b = ffi.new("size_t*"), ffi.new("size_t*")
    rmm.rmmGetInfo(free_b, total_b, ffi.cast("cudaStream_t", 0))
    assert 0 < free_b[0] <= total_b[0] and total_b[0] > (100 << 30)   # a B200 has 180 GB
    p = _ptr(rmm)
    rmm.rmmAlloc(p, 1 << 16, ffi.cast("cudaStream_t", 0))
    off = ffi.new("offset_t*")
    rmm.rmmGetAllocationOffset(off, p[0], ffi.cast("cudaStream_t", 0))
    assert off[0] == 0                                         # every rmm block is its own allocation here
    with pytest.raises(Exception):
        rmm.rmmGetAllocationOffset(off, ffi.NULL, ffi.cast("cudaStream_t", 0))
    rmm.rmmFree(p[0], ffi.cast("cudaStream_t", 0))


def test_write_log_and_get_log(rmm, tmp_path):
    ffi = rmm._ffi
    p = _ptr(rmm)
    rmm.rmmAlloc(p, 12345, ffi.cast("cudaStream_t", 0))
    rmm.rmmFree(p[0], ffi.cast("cudaStream_t", 0))
    path = str(tmp_path / "rmm.csv")
    rmm.rmmWriteLog(path.encode())
    text = open(path).read()
    assert text.startswith(HEADER) and ",12345," in text
    size = rmm._api.rmmLogSize()
    buf = ffi.new("char[]", size + 1)
    rmm.rmmGetLog(buf, size)
    assert ffi.string(buf, size).decode() == text
    with pytest.raises(Exception) as e:
        rmm.rmmWriteLog(os.path.join(str(tmp_path), "no", "such", "dir", "x.csv").encode())
    assert "RMM_ERROR_IO" in str(e.value)


def test_pool_keeps_blocks_and_trims(rmm):
    from libgdf_b200.librmm_cffi import librmm_config
    ffi = rmm._ffi
    p = _ptr(rmm)
    rmm.rmmAlloc(p, 64 << 20, ffi.cast("cudaStream_t", 0))
    first = int(ffi.cast("uintptr_t", p[0]))
    rmm.rmmFree(p[0], ffi.cast("cudaStream_t", 0))
    cached = rmm._api.rmmxPoolCachedBytes()
    if librmm_config.use_pool_allocator:
        assert cached >= (64 << 20)
        rmm.rmmAlloc(p, 64 << 20, ffi.cast("cudaStream_t", 0))
        assert int(ffi.cast("uintptr_t", p[0])) == first       # warm request: same block, no driver call
        rmm.rmmFree(p[0], ffi.cast("cudaStream_t", 0))
        rmm._api.rmmxTrimPool()
        assert rmm._api.rmmxPoolCachedBytes() == 0
    else:
        assert cached == 0


def test_library_scratch_cache_can_be_inspected_and_trimmed():
    from libgdf_b200.libgdf_cffi import libgdf_api
    import gpu_utils as G
    build = np.random.permutation(1_500_000).astype(np.int64)
    probe = np.random.randint(0, 1_500_000, 2_000_000).astype(np.int64)
    G.join("inner", [probe], [build])
    assert libgdf_api.gdfx_scratch_cached_bytes() > 0          # partition buffers and tables are kept for the next call
    released = libgdf_api.gdfx_trim_scratch()
    assert released > 0 and libgdf_api.gdfx_scratch_cached_bytes() == 0
    G.join("inner", [probe], [build])                          # and the library works after a trim


class _Alias(object):
    def __init__(self, address, nelem):
        self.__cuda_array_interface__ = {"shape": (nelem,), "typestr": "<i4", "data": (address, False), "version": 2,
                                         "strides": None}
