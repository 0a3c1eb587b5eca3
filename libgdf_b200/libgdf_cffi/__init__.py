"""ABI-mode cffi binding of libgdf.so - same import surface as the reference package
(reference: libgdf/python/libgdf_cffi/__init__.py:6-31): ``ffi``, ``libgdf``, ``GDFError``."""
import cffi

from .. import lib_path
from .._cdef import header_cdef
from .wrapper import GDFError, _libgdf_wrapper

ffi = cffi.FFI()
ffi.cdef(header_cdef("gdf/cffi/types.h", "gdf/cffi/functions.h", "gdf_b200_ext.h"))

# librmm.so is found through libgdf.so's $ORIGIN rpath.
libgdf_api = ffi.dlopen(lib_path("libgdf.so"))
libgdf = _libgdf_wrapper(ffi, libgdf_api)

__all__ = ["ffi", "libgdf", "libgdf_api", "GDFError"]
