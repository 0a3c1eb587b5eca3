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
