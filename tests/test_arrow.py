"""Arrow adapters (libgdf_b200/arrow.py): the reference's IPC parser contract
(python/tests/test_ipc.py:52-150 - schema JSON, layout JSON with data/null buffer offsets into the payload, data offset)
and the pyarrow / __cuda_array_interface__ column adapters.  Host-side logic runs without a GPU (torch CPU tensors stand
in for device memory: gdf_column_view only fills the struct); the -m gpu tests run gdf_* kernels on the adapted columns."""
import json

import numpy as np
import pyarrow as pa
import pytest
import torch

from libgdf_b200 import arrow as A
from libgdf_b200 import columns as C
from libgdf_b200.libgdf_cffi import ffi, libgdf


def make_batch(n=30, with_nulls=True):
    rng = np.random.RandomState(1234)
    idx = pa.array(np.arange(n, dtype=np.int32))
    w = rng.uniform(0, 1, n)
    mask = rng.rand(n) < 0.25 if with_nulls else np.zeros(n, bool)
    weight = pa.array(w, mask=mask)
    big = pa.array(rng.randint(-2 ** 40, 2 ** 40, n), type=pa.int64(), mask=(rng.rand(n) < 0.1) if with_nulls else None)
    day = pa.array(rng.randint(0, 20000, n).astype(np.int32), type=pa.date32())
    return pa.RecordBatch.from_arrays([idx, weight, big, day], ["idx", "weight", "big", "day"]), w, mask


def test_ipc_layout_matches_the_reference_contract():
    batch, w, mask = make_batch()
    schema_bytes = batch.schema.serialize().to_pybytes()
    payload = batch.serialize().to_pybytes()
    p = A.IpcParser(schema_bytes)
    sj = json.loads(p.schema_json())
    assert [f["name"] for f in sj["schema"]["fields"]] == ["idx", "weight", "big", "day"]      # test_ipc.py:112-115
    p.open_recordbatches(payload)
    rb = json.loads(p.layout_json())
    off = p.data_offset()
    assert [n["dtype"]["name"] for n in rb] == ["INT32", "DOUBLE", "INT64", "DATE32"]            # test_ipc.py:124,136,148
    assert [n["dtype"]["bitwidth"] for n in rb] == [32, 64, 64, 32]
    assert all(n["length"] == 30 for n in rb)
    body = np.frombuffer(payload, np.uint8)[off:]
    # the reference test reads every column straight out of the payload with these offsets (test_ipc.py:120-150)
    i0 = rb[0]["data_buffer"]
    np.testing.assert_array_equal(body[i0["offset"]:][:i0["length"]].view(np.int32)[:30], np.arange(30, dtype=np.int32))
    w0 = rb[1]["data_buffer"]
    got = body[w0["offset"]:][:w0["length"]].view(np.float64)[:30]
    np.testing.assert_array_equal(got[~mask], w[~mask])
    assert rb[1]["null_count"] == int(mask.sum())
    nb = rb[1]["null_buffer"]
    bits = np.unpackbits(body[nb["offset"]:][:nb["length"]], bitorder="little")[:30]
    np.testing.assert_array_equal(bits.astype(bool), ~mask)                                       # LSB-first, 1 = valid
    with pytest.raises(A.IpcParseError):
        p.open_recordbatches(payload)                                                             # "cannot open more than once"


def test_ipc_stream_with_schema_message_and_errors():
    batch, _, _ = make_batch(with_nulls=False)
    sink = pa.BufferOutputStream()
    with pa.ipc.new_stream(sink, batch.schema) as wr:
        wr.write_batch(batch)
    stream = sink.getvalue().to_pybytes()
    p = A.IpcParser(batch.schema.serialize().to_pybytes())
    p.open_recordbatches(stream)                         # schema message is skipped, end-of-stream marker honoured
    cols = p.columns(device="cpu")
    np.testing.assert_array_equal(cols["big"].data.numpy(), batch.column(2).to_numpy())
    assert cols["idx"].valid is None and cols["day"].dtype_name == "GDF_DATE32"
    with pytest.raises(A.IpcParseError):
        A.IpcParser(b"not a schema")
    q = A.IpcParser(batch.schema.serialize().to_pybytes())
    with pytest.raises(A.IpcParseError):
        q.layout_json()                                  # before open_recordbatches
    with pytest.raises(A.IpcParseError):
        q.open_recordbatches(stream[:40])                # truncated
    strings = pa.RecordBatch.from_arrays([pa.array(["a", "b"])], ["s"])
    with pytest.raises(A.IpcParseError):
        A.IpcParser(strings.schema.serialize().to_pybytes())   # no gdf_dtype for strings


def test_ipc_columns_alias_the_payload_and_carry_the_mask():
    batch, w, mask = make_batch(n=1003)
    p = A.IpcParser(batch.schema.serialize().to_pybytes())
    p.open_recordbatches(np.frombuffer(batch.serialize().to_pybytes(), np.uint8))
    cols = p.columns(device="cpu")
    c = cols["weight"]
    assert c.cdata.size == 1003 and c.cdata.null_count == int(mask.sum()) and c.dtype_name == "GDF_FLOAT64"
    np.testing.assert_array_equal(np.unpackbits(c.valid.numpy(), bitorder="little")[:1003].astype(bool), ~mask)
    base = cols.payload
    assert c.data.data_ptr() >= base.data_ptr() and c.data.data_ptr() < base.data_ptr() + base.numel()   # a view, not a copy


@pytest.mark.parametrize("n", [0, 1, 8, 77, 1000])
def test_from_arrow_to_arrow_round_trip(n):
    rng = np.random.RandomState(n)
    for t, np_t in ((pa.int8(), np.int8), (pa.int16(), np.int16), (pa.int32(), np.int32), (pa.int64(), np.int64),
                    (pa.float32(), np.float32), (pa.float64(), np.float64), (pa.date32(), np.int32), (pa.date64(), np.int64),
                    (pa.timestamp("ms"), np.int64)):
        vals = rng.randint(-100, 100, n).astype(np_t)
        if t == pa.date64():
            vals = vals.astype(np.int64) * 86400000
        mask = rng.rand(n) < 0.3
        arr = pa.array(vals, mask=mask).cast(t) if not pa.types.is_temporal(t) else pa.Array.from_buffers(
            t, n, [pa.array(vals, mask=mask).buffers()[0], pa.py_buffer(vals.tobytes())], null_count=int(mask.sum()))
        col = A.from_arrow(arr, device="cpu")
        assert col.cdata.size == n and col.cdata.null_count == int(mask.sum())
        back = A.to_arrow(col)
        assert back.equals(arr), (t, n)


def test_from_arrow_sliced_and_chunked_and_table():
    a = pa.array(np.arange(100, dtype=np.int64), mask=np.arange(100) % 7 == 0)
    sl = a.slice(13, 50)                                            # bit offset 13: re-based on the host
    col = A.from_arrow(sl, device="cpu")
    assert A.to_arrow(col).equals(pa.concat_arrays([sl]))
    ch = pa.chunked_array([a.slice(0, 40), a.slice(40)])
    assert A.to_arrow(A.from_arrow(ch, device="cpu")).equals(a)
    tbl = pa.table({"k": a, "v": pa.array(np.linspace(0, 1, 100))})
    cols = A.from_arrow(tbl, device="cpu")
    assert list(cols) == ["k", "v"] and cols["v"].valid is None and cols["k"].cdata.null_count == 15
    with pytest.raises(A.IpcParseError):
        A.from_arrow(pa.array(["x"]), device="cpu")


@pytest.mark.gpu
def test_device_resident_ipc_payload_feeds_the_kernels_zero_copy():
    """The reference's use of the parser: the record batch bytes are on the device; columns are views of them and go
    straight into gdf_* calls (here a masked gdf_sum_i64 and a gdf_add_f64)."""
    n = 100_003
    rng = np.random.RandomState(7)
    big = rng.randint(-2 ** 40, 2 ** 40, n)
    mask = rng.rand(n) < 0.2
    x, y = rng.rand(n), rng.rand(n)
    batch = pa.RecordBatch.from_arrays([pa.array(big, mask=mask), pa.array(x), pa.array(y)], ["big", "x", "y"])
    dev = torch.from_numpy(np.frombuffer(batch.serialize().to_pybytes(), np.uint8).copy()).cuda()
    p = A.IpcParser(batch.schema.serialize().to_pybytes())
    p.open_recordbatches(dev)
    cols = p.columns()
    lo, hi = dev.data_ptr(), dev.data_ptr() + dev.numel()
    assert all(lo <= c.data.data_ptr() < hi for c in cols.values()) and lo <= cols["big"].valid.data_ptr() < hi
    scratch = torch.zeros(int(libgdf.gdf_reduce_optimal_output_size()), dtype=torch.int64, device="cuda")
    libgdf.gdf_sum_i64(cols["big"].cdata, ffi.cast("int64_t*", scratch.data_ptr()), scratch.numel())
    assert int(scratch[0].item()) == int(big[~mask].sum())
    out = C.empty_column(n, torch.float64)
    libgdf.gdf_add_f64(cols["x"].cdata, cols["y"].cdata, out.cdata)
    np.testing.assert_array_equal(out.to_numpy(), x + y)


@pytest.mark.gpu
def test_cuda_array_interface_in_and_out():
    class Foreign(object):                                  # stands for a numba / cupy array
        def __init__(self, t):
            self.t = t
            self.__cuda_array_interface__ = t.__cuda_array_interface__
    a = torch.arange(1000, dtype=torch.int32, device="cuda")
    col = A.from_cuda_array(Foreign(a))
    assert int(ffi.cast("uintptr_t", col.cdata.data)) == a.data_ptr() and col.dtype_name == "GDF_INT32"
    out = C.empty_column(1000, torch.int32)
    libgdf.gdf_add_i32(col.cdata, col.cdata, out.cdata)
    view = A.CudaArrayView(out.cdata, np.int32, owner=out)
    back = torch.as_tensor(view, device="cuda")
    assert back.data_ptr() == out.data.data_ptr() and bool((back == 2 * a).all())
    # a join's library-owned outputs through the same export
    l = C.column(np.array([1, 2, 3, 4], np.int64))
    r = C.column(np.array([3, 1, 9], np.int64))
    ctx = ffi.new("gdf_context*")
    libgdf.gdf_context_view(ctx, 0, libgdf.GDF_HASH, 0, 0, 0)
    ol, orr = ffi.new("gdf_column*"), ffi.new("gdf_column*")
    libgdf.gdf_inner_join(C.column_array([l]), 1, ffi.new("int[]", [0]), C.column_array([r]), 1, ffi.new("int[]", [0]), 1, 0,
                          ffi.NULL, ol, orr, ctx)
    pairs = sorted(zip(torch.as_tensor(A.CudaArrayView(ol, np.int32), device="cuda").tolist(),
                       torch.as_tensor(A.CudaArrayView(orr, np.int32), device="cuda").tolist()))
    assert pairs == [(0, 1), (2, 0)]
    libgdf.gdf_column_free(ol), libgdf.gdf_column_free(orr)
