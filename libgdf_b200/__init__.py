"""libgdf_b200 - B200-native drop-in for the data-parallel hot path of gpuopenanalytics/libgdf.

The product is two C-ABI shared libraries built from ``csrc/`` (hand-written sm_100a CUDA):

* ``lib/libgdf.so``  - the ``gdf_*`` / ``gpu_*`` entry points declared in ``include/gdf/cffi/functions.h``
* ``lib/librmm.so``  - the ``rmm*`` allocator declared in ``include/memory.h``

and, on top of them, the same ABI-mode cffi bindings the reference ships
(reference: libgdf/python/libgdf_cffi, libgdf/python/librmm_cffi):

    from libgdf_b200.libgdf_cffi import ffi, libgdf, GDFError
    from libgdf_b200.librmm_cffi import librmm, librmm_config

``install_aliases()`` additionally registers the packages under the reference's own top-level names
(``libgdf_cffi``, ``librmm_cffi``) so unmodified reference-side code imports them.

There is no CPU fallback: if the shared libraries are missing the import raises, and every kernel
launch needs an sm_100a device.
"""
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
REPO_ROOT = os.path.dirname(_HERE)
LIB_DIR = os.path.join(_HERE, "lib")
INCLUDE_DIR = os.path.join(REPO_ROOT, "include")


def lib_path(name):
    """Absolute path of an in-tree shared library; raises if it has not been built."""
    path = os.path.join(LIB_DIR, name)
    if not os.path.isfile(path):
        raise ImportError(
            "%s is not built: run `make -C libgdf_b200/csrc` (or __graft_entry__.build()); "
            "libgdf_b200 has no CPU fallback" % path)
    return path


def install_aliases():
    """Expose the bindings under the reference's package names (libgdf_cffi / librmm_cffi)."""
    from . import libgdf_cffi, librmm_cffi
    sys.modules.setdefault("libgdf_cffi", libgdf_cffi)
    sys.modules.setdefault("librmm_cffi", librmm_cffi)
    return libgdf_cffi, librmm_cffi
