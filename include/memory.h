/* rmm* allocator ABI used by the gdf_* hot path for every library-owned output (join index columns,
 * result_cols) - reference: libgdf/include/memory.h:28-184.  Plain C, include-free (cdef()'d by
 * libgdf_b200/librmm_cffi).  Definitions: libgdf_b200/csrc/rmm.cpp (stream-ordered cudaMallocAsync pool).
 */
typedef struct CUstream_st *cudaStream_t;
typedef long int offset_t;

/* ref memory.h:33-43 */
typedef enum {
  RMM_SUCCESS = 0,
  RMM_ERROR_CUDA_ERROR,
  RMM_ERROR_INVALID_ARGUMENT,
  RMM_ERROR_NOT_INITIALIZED,
  RMM_ERROR_OUT_OF_MEMORY,
  RMM_ERROR_UNKNOWN,
  RMM_ERROR_IO,
  N_RMM_ERROR
} rmmError_t;

/* ref memory.h:45-56 */
typedef enum { CudaDefaultAllocation = 0, PoolAllocation } rmmAllocationMode_t;
typedef struct {
  rmmAllocationMode_t allocation_mode;
  size_t initial_pool_size;
  bool enable_logging;
} rmmOptions_t;

/* ref memory.h:65-184 */
rmmError_t rmmInitialize(rmmOptions_t *options);
rmmError_t rmmFinalize();
const char * rmmGetErrorString(rmmError_t errcode);
rmmError_t rmmAlloc(void **ptr, size_t size, cudaStream_t stream);
rmmError_t rmmRealloc(void **ptr, size_t new_size, cudaStream_t stream);
rmmError_t rmmFree(void *ptr, cudaStream_t stream);
rmmError_t rmmGetAllocationOffset(offset_t *offset, void *ptr, cudaStream_t stream);
rmmError_t rmmGetInfo(size_t *freeSize, size_t *totalSize, cudaStream_t stream);
rmmError_t rmmWriteLog(const char* filename);
size_t rmmLogSize();
rmmError_t rmmGetLog(char* buffer, size_t buffer_size);

/* ---- extensions (not in the reference): the PoolAllocation cache of this implementation keeps freed blocks for
 * reuse (csrc/block_cache.h); these two let a caller - and libgdf.so, before it reports out-of-memory for its own
 * scratch - hand the cached blocks back to the driver and see how much is parked. ---- */
void rmmxTrimPool(void);
size_t rmmxPoolCachedBytes(void);
