"""-m gpu: element-wise binary ops through the C ABI vs the numpy oracle.
Modelled on the reference's python/tests/test_binaryops.py (0-ulp equality, nelem=128, error codes
:203-270) and test_validity.py (masked add, :19-74), plus BASELINE config C1 (2 x 1M int32)."""
import numpy as np
import pytest
import torch

from oracle import np_oracle
from libgdf_b200 import columns as C
from libgdf_b200.libgdf_cffi import GDFError, ffi, libgdf
from gpu_utils import gen_rand, rand_mask

pytestmark = pytest.mark.gpu

ARITH = ["add", "sub", "mul", "floordiv"]
LOGIC = ["gt", "ge", "lt", "le", "eq", "ne"]
SUFFIX = {np.int8: "i8", np.int32: "i32", np.int64: "i64", np.float32: "f32", np.float64: "f64"}


def _run(fn_name, lhs, rhs, out_dtype, lvalid=None, rvalid=None, fill=-7):
    L, R = C.column(lhs, lvalid), C.column(rhs, rvalid)
    init = np.full(len(lhs), fill, dtype=out_dtype)
    O = C.column(init)
    getattr(libgdf, fn_name)(L.cdata, R.cdata, O.cdata)
    torch.cuda.synchronize()
    return O.to_numpy(), init


@pytest.mark.parametrize("nelem", [1, 2, 127, 128, 129, 1000, 100003])
@pytest.mark.parametrize("np_t", [np.int32, np.int64, np.float32, np.float64])
@pytest.mark.parametrize("op", ARITH)
def test_arith(op, np_t, nelem):
    lhs, rhs = gen_rand(np_t, nelem), gen_rand(np_t, nelem)
    if op == "floordiv":
        rhs[rhs == 0] = 1
    got, init = _run("gdf_%s_%s" % (op, SUFFIX[np_t]), lhs, rhs, np_t)
    want = np_oracle.binary_op(op, lhs, rhs, init)
    np.testing.assert_array_max_ulp(got, want, maxulp=0) if np.dtype(np_t).kind == "f" else np.testing.assert_array_equal(got, want)
    got2, _ = _run("gdf_%s_generic" % op, lhs, rhs, np_t)
    np.testing.assert_array_equal(got2, got)


@pytest.mark.parametrize("np_t", [np.float32, np.float64])
def test_div(np_t):
    lhs, rhs = gen_rand(np_t, 1000), gen_rand(np_t, 1000)
    got, init = _run("gdf_div_%s" % SUFFIX[np_t], lhs, rhs, np_t)
    np.testing.assert_array_max_ulp(got, np_oracle.binary_op("div", lhs, rhs, init), maxulp=0)


@pytest.mark.parametrize("np_t", [np.int8, np.int32, np.int64, np.float32, np.float64])
@pytest.mark.parametrize("op", LOGIC)
def test_logical(op, np_t):
    lhs, rhs = gen_rand(np_t, 1000, low=-5, high=5), gen_rand(np_t, 1000, low=-5, high=5)
    got, init = _run("gdf_%s_%s" % (op, SUFFIX[np_t]), lhs, rhs, np.int8)
    np.testing.assert_array_equal(got, np_oracle.binary_op(op, lhs, rhs, init))


@pytest.mark.parametrize("np_t", [np.int8, np.int32, np.int64])
@pytest.mark.parametrize("op", ["bitwise_and", "bitwise_or", "bitwise_xor"])
def test_bitwise(op, np_t):
    lhs, rhs = gen_rand(np_t, 777), gen_rand(np_t, 777)
    got, init = _run("gdf_%s_%s" % (op, SUFFIX[np_t]), lhs, rhs, np_t)
    np.testing.assert_array_equal(got, np_oracle.binary_op(op, lhs, rhs, init))


@pytest.mark.parametrize("nelem", [5, 128, 1001])
def test_masked_add_only_touches_valid_lanes(nelem):
    lhs, rhs = gen_rand(np.int32, nelem), gen_rand(np.int32, nelem)
    lm, _ = rand_mask(nelem)
    rm, _ = rand_mask(nelem)
    got, init = _run("gdf_add_i32", lhs, rhs, np.int32, lm, rm)
    np.testing.assert_array_equal(got, np_oracle.binary_op("add", lhs, rhs, init, lm, rm))
    got, init = _run("gdf_add_i32", lhs, rhs, np.int32, lm, None)
    np.testing.assert_array_equal(got, np_oracle.binary_op("add", lhs, rhs, init, lm, None))


def test_validity_and():
    n = 1000
    lm, _ = rand_mask(n)
    rm, _ = rand_mask(n)
    L, R = C.column(np.zeros(n, np.int32), lm), C.column(np.zeros(n, np.int32), rm)
    O = C.column(np.zeros(n, np.int32), np.zeros_like(lm))
    libgdf.gdf_validity_and(L.cdata, R.cdata, O.cdata)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(O.valid.cpu().numpy(), lm & rm)


def test_c1_add_1m_int32():
    """BASELINE config C1: gdf_add on two 1M-row int32 columns, uniform [-10000, 10000)."""
    lhs, rhs = gen_rand(np.int32, 1_000_000), gen_rand(np.int32, 1_000_000)
    got, _ = _run("gdf_add_generic", lhs, rhs, np.int32)
    np.testing.assert_array_equal(got, lhs + rhs)


def test_error_codes():
    a, b = C.column(np.zeros(8, np.int32)), C.column(np.zeros(9, np.int32))
    o = C.column(np.zeros(8, np.int32))
    with pytest.raises(GDFError) as e:
        libgdf.gdf_add_i32(a.cdata, b.cdata, o.cdata)
    assert e.value.errcode == "GDF_COLUMN_SIZE_MISMATCH"
    f = C.column(np.zeros(8, np.float32))
    with pytest.raises(GDFError) as e:
        libgdf.gdf_add_generic(a.cdata, f.cdata, o.cdata)
    assert e.value.errcode == "GDF_UNSUPPORTED_DTYPE"
    o64 = C.column(np.zeros(8, np.int64))
    with pytest.raises(GDFError) as e:
        libgdf.gdf_add_i32(a.cdata, a.cdata, o64.cdata)
    assert e.value.errcode == "GDF_UNSUPPORTED_DTYPE"
    e0 = C.column(np.zeros(0, np.int32))
    assert libgdf.gdf_add_i32(e0.cdata, e0.cdata, e0.cdata) is None   # empty -> success
