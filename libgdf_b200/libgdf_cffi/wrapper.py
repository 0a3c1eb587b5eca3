"""Error-code -> exception adapter around the raw cffi library object.

Same contract as the reference wrapper (reference: libgdf/python/libgdf_cffi/wrapper.py:1-52): every
function whose C return type is ``gdf_error`` raises ``GDFError(errname, msg)`` on a non-zero code;
a GDF_CUDA_ERROR is expanded with the CUDA error name and string."""


class GDFError(Exception):
    def __init__(self, errcode, msg):
        self.errcode = errcode
        super(GDFError, self).__init__(msg)


class _libgdf_wrapper(object):
    def __init__(self, ffi, api):
        self._ffi = ffi
        self._api = api
        self._cached = {}

    def __getattr__(self, name):
        cached = self._cached.get(name)
        if cached is not None:
            return cached
        fn = getattr(self._api, name)
        if callable(fn) and self._returns_gdf_error(fn):
            fn = self._checked(fn, name)
        self._cached[name] = fn
        return fn

    def _returns_gdf_error(self, fn):
        try:
            return self._ffi.typeof(fn).result.cname == "gdf_error"
        except TypeError:
            return False

    def _checked(self, fn, name):
        def wrap(*args):
            errcode = fn(*args)
            if errcode != self._api.GDF_SUCCESS:
                errname, msg = self._get_error_msg(errcode)
                raise GDFError(errname, msg)
        wrap.__name__ = name
        return wrap

    def _ffi_str(self, strptr):
        return self._ffi.string(strptr).decode("ascii")

    def _get_error_msg(self, errcode):
        if errcode == self._api.GDF_CUDA_ERROR:
            cudaerr = self._api.gdf_cuda_last_error()
            errname = self._ffi_str(self._api.gdf_cuda_error_name(cudaerr))
            details = self._ffi_str(self._api.gdf_cuda_error_string(cudaerr))
            return errname, "CUDA ERROR. {}: {}".format(errname, details)
        errname = self._ffi_str(self._api.gdf_error_get_name(errcode))
        return errname, errname
