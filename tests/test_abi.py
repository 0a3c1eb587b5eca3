"""CPU tests of the drop-in boundary: the shared libraries load without a GPU, export every symbol
the public headers declare, keep the reference's struct layout, and the host-only entry points work."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return re.findall(r"\b((?:gdf|gpu|rmm|get)_?\w+)\s*\(", text)


def test_libgdf_exports_every_declared_symbol():
    import libgdf_b200
    lib = ctypes.CDLL(libgdf_b200.lib_path("libgdf.so"))
    names = _declared("gdf/cffi/functions.h")
    assert len(names) > 120
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_libgdf_exports_every_extension_symbol():
    """include/gdf_b200_ext.h: measurement hooks and the multi-GPU layer's entry points (not in the reference)."""
    import libgdf_b200
    lib = ctypes.CDLL(libgdf_b200.lib_path("libgdf.so"))
    names = _declared("gdf_b200_ext.h")
    assert {"gdfx_profile_enable", "gdfx_partition_pairs", "gdfx_join_pairs", "gdfx_partition_scatter_peer",
            "gdfx_peer_alloc", "gdfx_rows_valid_to_bytes"} <= set(names)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_librmm_exports_every_declared_symbol():
    import libgdf_b200
    lib = ctypes.CDLL(libgdf_b200.lib_path("librmm.so"))
    names = [n for n in _declared("memory.h") if n.startswith("rmm")]
    assert len([n for n in names if not n.startswith("rmmx")]) == 11       # the reference's ABI (memory.h:65-184)
    assert {"rmmxTrimPool", "rmmxPoolCachedBytes"} <= set(names)            # + this implementation's extensions
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_struct_layout_matches_reference(gdf):
    ffi, libgdf = gdf
    # reference types.h:84-92 (offsets verified in SURVEY.md section 8 a1)
    assert ffi.sizeof("gdf_column") == 56 == libgdf.gdf_column_sizeof()
    for field, off in (("data", 0), ("valid", 8), ("size", 16), ("dtype", 24), ("null_count", 32),
                       ("dtype_info", 40), ("col_name", 48)):
        assert ffi.offsetof("gdf_column", field) == off
    assert ffi.sizeof("gdf_context") == 20
    assert libgdf.GDF_INT64 == 4 and libgdf.GDF_FLOAT64 == 6 and libgdf.N_GDF_TYPES == 12
    assert libgdf.GDF_HASH == 1 and libgdf.GDF_COUNT == 4 and libgdf.N_GDF_ERRORS == 23


def test_error_names(gdf):
    ffi, libgdf = gdf
    assert ffi.string(libgdf.gdf_error_get_name(libgdf.GDF_SUCCESS)) == b"GDF_SUCCESS"
    assert ffi.string(libgdf.gdf_error_get_name(libgdf.GDF_JOIN_DTYPE_MISMATCH)) == b"GDF_JOIN_DTYPE_MISMATCH"
    assert ffi.string(libgdf.gdf_error_get_name(libgdf.GDF_NULL_NVTX_NAME)) == b"GDF_NULL_NVTX_NAME"
    assert b"Unknown error" in ffi.string(libgdf.gdf_error_get_name(9999))


def test_column_and_context_views(gdf):
    ffi, libgdf = gdf
    col = ffi.new("gdf_column*")
    libgdf.gdf_column_view_augmented(col, ffi.cast("void*", 0x1000), ffi.NULL, 77, libgdf.GDF_INT32, 5)
    assert (int(ffi.cast("uintptr_t", col.data)), col.size, col.dtype, col.null_count) == (0x1000, 77, libgdf.GDF_INT32, 5)
    libgdf.gdf_column_view(col, ffi.NULL, ffi.NULL, 3, libgdf.GDF_FLOAT64)
    assert col.null_count == 0 and col.size == 3
    width = ffi.new("int*")
    libgdf.get_column_byte_width(col, width)
    assert width[0] == 8
    ctx = ffi.new("gdf_context*")
    libgdf.gdf_context_view(ctx, 0, libgdf.GDF_HASH, 0, 1, 0)
    assert (ctx.flag_method, ctx.flag_sort_result) == (libgdf.GDF_HASH, 1)


def test_wrapper_raises_gdferror(gdf):
    ffi, libgdf = gdf
    from libgdf_b200.libgdf_cffi import GDFError
    col = ffi.new("gdf_column*")
    libgdf.gdf_column_view(col, ffi.NULL, ffi.NULL, 0, libgdf.GDF_STRING)
    width = ffi.new("int*")
    with pytest.raises(GDFError) as exc:
        libgdf.get_column_byte_width(col, width)
    assert exc.value.errcode == "GDF_UNSUPPORTED_DTYPE"
    assert libgdf.gdf_nvtx_range_push(b"range", libgdf.GDF_GREEN) is None
    assert libgdf.gdf_nvtx_range_pop() is None
    with pytest.raises(GDFError):
        libgdf.gdf_nvtx_range_push(ffi.NULL, libgdf.GDF_GREEN)


def test_argument_errors_need_no_gpu(gdf):
    """Validation happens on the host before any launch (reference: joining.cu:290-301, sqls_ops.cu:1095-1106)."""
    ffi, libgdf = gdf
    from libgdf_b200.libgdf_cffi import GDFError
    ctx = ffi.new("gdf_context*")
    libgdf.gdf_context_view(ctx, 0, libgdf.GDF_HASH, 0, 0, 0)
    out = ffi.new("gdf_column*")
    with pytest.raises(GDFError) as exc:
        libgdf.gdf_inner_join(ffi.NULL, 0, ffi.NULL, ffi.NULL, 0, ffi.NULL, 1, 0, ffi.NULL, out, out, ctx)
    assert exc.value.errcode == "GDF_DATASET_EMPTY"
    with pytest.raises(GDFError) as exc:
        libgdf.gdf_group_by_sum(0, ffi.NULL, ffi.NULL, ffi.NULL, ffi.NULL, ffi.NULL, ctx)
    assert exc.value.errcode == "GDF_DATASET_EMPTY"
    with pytest.raises(GDFError) as exc:
        libgdf.gdf_hash_partition(0, ffi.NULL, ffi.NULL, 0, 0, ffi.NULL, ffi.NULL, libgdf.GDF_HASH_MURMUR3)
    assert exc.value.errcode == "GDF_INVALID_API_CALL"
    assert libgdf.gdf_reduce_optimal_output_size() == 128


def test_rmm_error_strings():
    from libgdf_b200.librmm_cffi import ffi, librmm_api
    assert ffi.string(librmm_api.rmmGetErrorString(librmm_api.RMM_SUCCESS)) == b"RMM_SUCCESS"
    assert ffi.string(librmm_api.rmmGetErrorString(librmm_api.RMM_ERROR_OUT_OF_MEMORY)) == b"RMM_ERROR_OUT_OF_MEMORY"


def test_product_never_imports_oracle():
    """The product path must not route through the CPU oracle (or any CPU fallback)."""
    pkg = os.path.join(ROOT, "libgdf_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "liboracle" not in text, f


def test_reference_package_names_resolve_after_install_aliases():
    """INTEGRATION.md route (a): unmodified reference-side imports find the drop-in bindings."""
    import libgdf_b200
    libgdf_b200.install_aliases()
    from libgdf_cffi import ffi, libgdf, GDFError          # noqa: F401  (the reference's import line)
    from librmm_cffi import librmm, librmm_config           # noqa: F401
    assert libgdf.gdf_column_sizeof() == ffi.sizeof("gdf_column") == 56
    for name in ("initialize", "finalize", "to_device", "device_array", "device_array_like", "csv_log"):
        assert hasattr(librmm, name)
