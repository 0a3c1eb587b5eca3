"""Run-time configuration read by ``librmm.initialize()`` (reference:
libgdf/python/librmm_cffi/librmm_config.py).  Set these before importing ``librmm_cffi``."""

# False: cudaMalloc/cudaFree per allocation.  True: caching pool (csrc/block_cache.h): freed blocks are kept and
# reused, stream-aware like the reference's cnmem pool; librmm.finalize() / rmmxTrimPool() give them back.
use_pool_allocator = False

# Bytes reserved up front when the pool allocator is on; 0 = grow on demand.
initial_pool_size = 0

# Record every alloc/realloc/free (retrievable with librmm.csv_log()).
enable_logging = False
