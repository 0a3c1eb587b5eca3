"""-m gpu: columnar reductions through the C ABI vs the numpy oracle.  Sizes and dtypes follow the
reference's python/tests/test_reductions.py (:29-50,84-105,115-183): integers bit-exact, floats to the
reference's own tolerance (5 decimals relative here, stated below)."""
import numpy as np
import pytest
import torch

from oracle import np_oracle
import oracle
from libgdf_b200 import columns as C
from libgdf_b200.libgdf_cffi import ffi, libgdf
from gpu_utils import gen_rand, rand_mask

pytestmark = pytest.mark.gpu
SUFFIX = {np.int8: "i8", np.int32: "i32", np.int64: "i64", np.float32: "f32", np.float64: "f64"}
RTOL = {np.float32: 1e-4, np.float64: 1e-9}   # floating sums differ from the oracle only by order


def _reduce(name, data, valid=None, out_size=None):
    col = C.column(data, valid)
    out_size = out_size or libgdf.gdf_reduce_optimal_output_size()
    out = torch.zeros(out_size, dtype=col.data.dtype, device="cuda")
    fn = getattr(libgdf, "gdf_%s_%s" % (name, SUFFIX[data.dtype.type]))
    ptr = ffi.cast({np.int8: "int8_t*", np.int32: "int32_t*", np.int64: "int64_t*", np.float32: "float*",
                    np.float64: "double*"}[data.dtype.type], out.data_ptr())
    fn(col.cdata, ptr, out_size)
    torch.cuda.synchronize()
    return out[0].cpu().numpy()


def _check(got, want, np_t):
    if np.dtype(np_t).kind == "f":
        np.testing.assert_allclose(got, want, rtol=RTOL[np_t], atol=RTOL[np_t])
    else:
        assert got == want


@pytest.mark.parametrize("nelem", [1, 2, 3, 127, 128, 129, 200, 10000, 1_000_003])
@pytest.mark.parametrize("np_t", [np.float64, np.float32, np.int64, np.int32, np.int8])
@pytest.mark.parametrize("name", ["sum", "min", "max"])
def test_reduce(name, np_t, nelem):
    data = gen_rand(np_t, nelem)
    _check(_reduce(name, data), np_oracle.reduce(name, data), np_t)


@pytest.mark.parametrize("nelem", [1, 3, 129, 10000])
@pytest.mark.parametrize("np_t", [np.float64, np.float32, np.int64, np.int32, np.int8])
def test_product(np_t, nelem):
    data = gen_rand(np_t, nelem, low=-3, high=4)
    if np.dtype(np_t).kind == "f":
        data = (1.0 + data * 1e-3).astype(np_t)       # keep the product away from 0 / inf
        np.testing.assert_allclose(_reduce("product", data), np_oracle.reduce("product", data), rtol=1e-3)
    else:
        assert _reduce("product", data) == np_oracle.reduce("product", data)


@pytest.mark.parametrize("np_t", [np.float64, np.float32])
def test_sum_squared(np_t):
    data = gen_rand(np_t, 10000)
    _check(_reduce("sum_squared", data), np_oracle.reduce("sum_squared", data), np_t)


@pytest.mark.parametrize("nelem", [1, 7, 8, 9, 129, 10001, 1_000_003])
@pytest.mark.parametrize("np_t", [np.int64, np.int32, np.int8, np.float64])
@pytest.mark.parametrize("name", ["sum", "min", "max"])
def test_masked(name, np_t, nelem):
    data = gen_rand(np_t, nelem)
    mask, _ = rand_mask(nelem)
    _check(_reduce(name, data, mask), np_oracle.reduce(name, data, mask), np_t)


def test_all_null_and_empty_return_identity():
    data = gen_rand(np.int32, 100)
    assert _reduce("min", data, np.zeros(13, np.uint8)) == np.iinfo(np.int32).max
    assert _reduce("max", data, np.zeros(13, np.uint8)) == np.iinfo(np.int32).min
    assert _reduce("sum", np.zeros(0, np.int64)) == 0


def test_single_element_scratch_and_generic():
    data = gen_rand(np.int64, 5000)
    assert _reduce("sum", data, out_size=1) == data.sum()
    col = C.column(data)
    out = torch.zeros(128, dtype=torch.int64, device="cuda")
    libgdf.gdf_sum_generic(col.cdata, ffi.cast("void*", out.data_ptr()), 128)
    assert int(out[0].item()) == int(data.sum()) == oracle.sum_i64(data)


def test_unaligned_pointer():
    base = torch.as_tensor(gen_rand(np.int64, 4097)).cuda()
    view = base[1:]                      # 8-byte aligned only
    col = C.Column(view)
    out = torch.zeros(128, dtype=torch.int64, device="cuda")
    libgdf.gdf_sum_i64(col.cdata, ffi.cast("int64_t*", out.data_ptr()), 128)
    assert int(out[0].item()) == int(view.sum().item())


def test_count_nonzero_mask():
    for n in (1, 7, 8, 9, 1000, 100003):
        mask, bits = rand_mask(n)
        m = torch.as_tensor(mask).cuda()
        cnt = ffi.new("int*")
        libgdf.gdf_count_nonzero_mask(ffi.cast("gdf_valid_type*", m.data_ptr()), n, cnt)
        assert cnt[0] == int(bits.sum())
