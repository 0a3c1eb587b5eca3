"""Python face of librmm.so.

Follows the reference wrapper's API (reference: libgdf/python/librmm_cffi/wrapper.py:73-232):
``initialize / finalize / csv_log / device_array / device_array_like / to_device / get_ipc_handle``.
The reference hands out numba device arrays; numba's CUDA driver layer is not usable in this image,
so device arrays are ``torch`` CUDA tensors that alias rmm-owned memory through
``__cuda_array_interface__`` (torch is the device-buffer plumbing of this repo) and return the block to
rmm when garbage-collected."""
import weakref

import numpy as np

from . import librmm_config as rmm_cfg


class RMMError(Exception):
    def __init__(self, errcode, msg):
        self.errcode = errcode
        super(RMMError, self).__init__(msg)


class _RMMBlock(object):
    """Owner of one rmmAlloc'd block, exported through the CUDA array interface."""

    def __init__(self, wrapper, nbytes, shape, dtype, stream):
        self._wrapper = wrapper
        self._stream = stream
        ptr = wrapper._ffi.new("void **")
        wrapper.rmmAlloc(ptr, max(int(nbytes), 1), wrapper._ffi.cast("cudaStream_t", stream))
        self.address = int(wrapper._ffi.cast("uintptr_t", ptr[0]))
        self.__cuda_array_interface__ = {
            "shape": tuple(shape), "typestr": np.dtype(dtype).str, "data": (self.address, False),
            "version": 2, "strides": None,
        }
        weakref.finalize(self, _RMMBlock._release, wrapper, self.address, stream)

    @staticmethod
    def _release(wrapper, address, stream):
        try:
            wrapper._api.rmmFree(wrapper._ffi.cast("void*", address), wrapper._ffi.cast("cudaStream_t", stream))
        except Exception:  # interpreter shutdown
            pass


class _RMMWrapper(object):
    def __init__(self, ffi, api):
        self._ffi = ffi
        self._api = api
        self._cached = {}

    def __getattr__(self, name):
        cached = self._cached.get(name)
        if cached is not None:
            return cached
        fn = getattr(self._api, name)
        try:
            checked = self._ffi.typeof(fn).result.cname == "rmmError_t"
        except TypeError:
            checked = False
        if checked:
            raw = fn

            def fn(*args):
                errcode = raw(*args)
                if errcode != self._api.RMM_SUCCESS:
                    errname = self._ffi.string(self._api.rmmGetErrorString(errcode)).decode("ascii")
                    raise RMMError(errname, "RMM error encountered: {}".format(errname))
            fn.__name__ = name
        self._cached[name] = fn
        return fn

    # ---- reference API ---------------------------------------------------------------------
    def initialize(self):
        opts = self._ffi.new("rmmOptions_t *")
        opts.allocation_mode = (self._api.PoolAllocation if rmm_cfg.use_pool_allocator
                                else self._api.CudaDefaultAllocation)
        opts.initial_pool_size = rmm_cfg.initial_pool_size
        opts.enable_logging = rmm_cfg.enable_logging
        return self.rmmInitialize(opts)

    def finalize(self):
        return self.rmmFinalize()

    def csv_log(self):
        size = self._api.rmmLogSize()
        buf = self._ffi.new("char[]", size + 1)
        self.rmmGetLog(buf, size)
        return self._ffi.string(buf, size).decode("utf-8")

    def device_array(self, shape, dtype=np.float64, strides=None, order="C", stream=0):
        import torch
        if isinstance(shape, int):
            shape = (shape,)
        dtype = np.dtype(dtype)
        nelem = int(np.prod(shape)) if len(shape) else 1
        block = _RMMBlock(self, nelem * dtype.itemsize, shape, dtype, stream)
        if nelem == 0:
            return torch.empty(tuple(shape), dtype=getattr(torch, dtype.name), device="cuda")
        tensor = torch.as_tensor(block, device="cuda")
        tensor._rmm_block = block  # keep the allocation alive as long as the tensor
        return tensor

    def device_array_like(self, ary, stream=0):
        ary = np.asarray(ary) if not hasattr(ary, "shape") else ary
        shape = tuple(ary.shape) if len(ary.shape) else (1,)
        dtype = ary.dtype if isinstance(ary.dtype, np.dtype) else np.dtype(str(ary.dtype).replace("torch.", ""))
        return self.device_array(shape, dtype, stream=stream)

    def to_device(self, ary, stream=0, copy=True, to=None):
        import torch
        ary = np.ascontiguousarray(ary)
        if to is None:
            to = self.device_array_like(ary, stream=stream)
            copy = True
        if copy and ary.size:
            to.copy_(torch.from_numpy(ary.reshape(to.shape)))
        return to

    def get_ipc_handle(self, ary, stream=0):
        """(cudaIpc handle bytes, offset) of a device array; offset comes from rmmGetAllocationOffset."""
        offset = self._ffi.new("offset_t*")
        ptr = self._ffi.cast("void*", ary.data_ptr())
        self.rmmGetAllocationOffset(offset, ptr, self._ffi.cast("cudaStream_t", stream))
        storage = ary.untyped_storage()
        handle = storage._share_cuda_()
        return handle, int(offset[0])
