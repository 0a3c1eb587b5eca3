"""torch <-> gdf_column plumbing for callers of the C ABI from Python.

The reference's Python tests wrap numba device arrays into ``gdf_column`` structs with
``gdf_column_view`` (reference: libgdf/python/tests/utils.py:7-28).  numba's CUDA layer is unusable
in this image, so the same thing is done with torch CUDA tensors: PyTorch here is device memory and
streams only - every computation goes through ``libgdf.so``.
"""
import numpy as np
import torch

from .libgdf_cffi import ffi, libgdf

_TORCH_TO_GDF = {
    torch.int8: "GDF_INT8", torch.int16: "GDF_INT16", torch.int32: "GDF_INT32", torch.int64: "GDF_INT64",
    torch.float32: "GDF_FLOAT32", torch.float64: "GDF_FLOAT64",
}
_GDF_TO_TORCH = {
    "GDF_INT8": torch.int8, "GDF_INT16": torch.int16, "GDF_INT32": torch.int32, "GDF_INT64": torch.int64,
    "GDF_FLOAT32": torch.float32, "GDF_FLOAT64": torch.float64, "GDF_DATE32": torch.int32,
    "GDF_DATE64": torch.int64, "GDF_TIMESTAMP": torch.int64,
}


def gdf_dtype_of(tensor, override=None):
    name = override or _TORCH_TO_GDF[tensor.dtype]
    return getattr(libgdf, name), name


def valid_nbytes(rows):
    return (rows + 7) // 8


class Column(object):
    """A ``gdf_column*`` plus the torch tensors that own its device buffers."""

    def __init__(self, data, valid=None, dtype=None, null_count=None, api=None):
        api = api or libgdf
        self.data = data
        self.valid = valid
        self.cdata = ffi.new("gdf_column*")
        code, self.dtype_name = gdf_dtype_of(data, dtype)
        data_ptr = ffi.cast("void*", data.data_ptr()) if data.numel() else ffi.NULL
        valid_ptr = ffi.cast("gdf_valid_type*", valid.data_ptr()) if valid is not None else ffi.NULL
        if null_count is None and valid is not None:
            null_count = data.numel() - count_valid(valid, data.numel())
        api.gdf_column_view_augmented(self.cdata, data_ptr, valid_ptr, data.numel(), code, null_count or 0)

    @property
    def size(self):
        return int(self.cdata.size)

    def to_numpy(self):
        return self.data[: self.size].cpu().numpy()


def count_valid(valid, rows):
    bits = np.unpackbits(valid.cpu().numpy(), bitorder="little")[:rows]
    return int(bits.sum())


def column(array, valid=None, dtype=None, device="cuda", api=None):
    """Build a Column from a numpy array / torch tensor (+ optional packed validity bytes)."""
    data = torch.as_tensor(np.ascontiguousarray(array) if isinstance(array, np.ndarray) else array).to(device)
    v = None
    if valid is not None:
        v = torch.as_tensor(np.ascontiguousarray(valid, dtype=np.uint8) if isinstance(valid, np.ndarray) else valid).to(device)
    return Column(data, v, dtype=dtype, api=api)


def empty_column(rows, torch_dtype, with_valid=False, dtype=None, device="cuda", api=None):
    data = torch.empty(rows, dtype=torch_dtype, device=device)
    v = torch.zeros(valid_nbytes(rows), dtype=torch.uint8, device=device) if with_valid else None
    return Column(data, v, dtype=dtype, null_count=0, api=api)


def column_array(cols):
    """``gdf_column*[]`` from Columns; keeps them alive on the returned object."""
    arr = ffi.new("gdf_column*[]", [c.cdata for c in cols])
    return arr


def struct_array(cols):
    """Host array of ``gdf_column`` STRUCTS (what gdf_filter takes)."""
    arr = ffi.new("gdf_column[]", len(cols))
    for i, c in enumerate(cols):
        arr[i] = c.cdata[0]
    return arr


def library_owned_to_torch(cdata, api=None):
    """Copy a library-allocated column (join output) into a torch tensor and free it with
    gdf_column_free, the way a reference caller would (reference: src/column.cpp:222-227)."""
    api = api or libgdf
    n = int(cdata.size)
    out = torch.empty(n, dtype=torch.int32, device="cuda")
    if n:
        src = _alias(int(ffi.cast("uintptr_t", cdata.data)), n, np.int32)
        out.copy_(src)
        torch.cuda.synchronize()
    if cdata.data != ffi.NULL:
        api.gdf_column_free(cdata)
    return out


class _OwnedAlias(object):
    """__cuda_array_interface__ view of a library-allocated column that frees it (gdf_column_free -> rmmFree, reference
    src/column.cpp:222-227) when the last tensor built on it goes away: torch keeps the producer object alive."""

    def __init__(self, cdata, np_dtype, api):
        self._api = api
        self._col = ffi.new("gdf_column*")
        self._col[0] = cdata[0]          # own a copy of the struct: the caller's gdf_column may be reused
        n = int(cdata.size)
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": np.dtype(np_dtype).str,
                                         "data": (int(ffi.cast("uintptr_t", cdata.data)), False), "version": 2, "strides": None}

    def __del__(self):
        try:
            if self._col.data != ffi.NULL:
                self._api.gdf_column_free(self._col)
        except Exception:      # interpreter shutdown
            pass


def library_owned_view(cdata, api=None):
    """Zero-copy torch view of a library-allocated GDF_INT32 column (join output); the memory is handed back with
    gdf_column_free when the tensor dies.  No copy, no synchronisation (library_owned_to_torch does both)."""
    api = api or libgdf
    n = int(cdata.size)
    if n == 0:
        if cdata.data != ffi.NULL:
            api.gdf_column_free(cdata)
        return torch.empty(0, dtype=torch.int32, device="cuda")
    return torch.as_tensor(_OwnedAlias(cdata, np.int32, api), device="cuda")


class _Alias(object):
    def __init__(self, address, shape, dtype):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": np.dtype(dtype).str,
                                         "data": (address, False), "version": 2, "strides": None}


def _alias(address, nelem, dtype):
    """Zero-copy torch view of foreign device memory."""
    return torch.as_tensor(_Alias(address, (nelem,), dtype), device="cuda")


def alias_column_data(cdata, np_dtype):
    """Zero-copy torch view of a gdf_column's data buffer (caller guarantees lifetime)."""
    n = int(cdata.size)
    if n == 0:
        return torch.empty(0, dtype=getattr(torch, np.dtype(np_dtype).name), device="cuda")
    return _alias(int(ffi.cast("uintptr_t", cdata.data)), n, np_dtype)
