"""ABI-mode cffi binding of librmm.so - same import surface as the reference package
(reference: libgdf/python/librmm_cffi/__init__.py:16-50): ``librmm``, ``librmm_config``; importing the
package initialises the memory manager and registers ``finalize`` at interpreter exit."""
import atexit

import cffi

from .. import lib_path
from .._cdef import header_cdef
from . import librmm_config
from .wrapper import RMMError, _RMMWrapper

ffi = cffi.FFI()
ffi.cdef(header_cdef("memory.h"))

librmm_api = ffi.dlopen(lib_path("librmm.so"))
librmm = _RMMWrapper(ffi, librmm_api)
librmm.initialize()
atexit.register(librmm.finalize)

__all__ = ["ffi", "librmm", "librmm_api", "librmm_config", "RMMError"]
