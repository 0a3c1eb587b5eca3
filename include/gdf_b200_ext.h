/* Extensions that are NOT part of the reference ABI (plain C, include-free, cdef()'d after types.h).
 * They exist for measurement and for the multi-GPU layer, which the reference does not have
 * (no NCCL/MPI/multi-device code anywhere in libgdf). */

/* Per-kernel device timing.  When enabled, every kernel launch of the library is bracketed by CUDA
 * events on the launching (legacy default) stream; gdfx_profile_report() synchronises, folds the
 * pending pairs into per-kernel {launches, total ms} and writes them as one JSON object into buf
 * (always NUL-terminated; returns the number of bytes the full report needs).  bench.py uses this to
 * measure the dominant kernel's launch duration live, inside the timed region. */
int gdfx_profile_enable(int on);
size_t gdfx_profile_report(char *buf, size_t capacity);

/* Multi-GPU layer helper (libgdf_b200/dist.py): indices[i] = payload[indices[i]] in place for every
 * non-negative entry (< payload_rows), -1 otherwise.  `indices` is a GDF_INT32 join output column;
 * `payload` is a device array of the global row ids that travelled with the exchanged keys. */
gdf_error gdfx_remap_indices(gdf_column *indices, const int32_t *payload, size_t payload_rows);
