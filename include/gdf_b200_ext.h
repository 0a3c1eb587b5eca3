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

/* Library-internal scratch memory (hash tables, partition buffers) is cached between calls so that warm calls
 * never enter the driver (csrc/block_cache.h).  The cache is bounded (default: half of the device's memory) and is
 * handed back automatically when one of the library's own allocations runs out of memory; a caller that shares
 * the GPU with other allocators can also empty it (returns the bytes released), inspect it, or bound it. */
size_t gdfx_trim_scratch(void);
size_t gdfx_scratch_cached_bytes(void);
void gdfx_set_scratch_limit(size_t bytes);

/* gdf_filter on one aligned column runs a persistent kernel whose chunk look-back needs forward progress of the CTAs
 * holding earlier chunks (csrc/select_stream.cuh).  mode 0 (default): static chunk dealing under a COOPERATIVE launch
 * (the driver starts the grid only once all of its CTAs fit), falling back to mode 1 when the device refuses one;
 * mode 1: chunks are taken through an atomic ticket, which needs no co-residency at all (about 20 % slower at 1e9
 * rows).  Returns the previous mode.  gdfx_debug_occupy_sms is a TEST helper: it parks `blocks` CTAs that fill an SM
 * each (1024 threads, 200 KB of shared memory) on a private non-blocking stream for about `microseconds`, so that a
 * test can run gdf_filter while a foreign kernel holds part of the GPU (tests/test_filter_gpu.py). */
int gdfx_set_select_dealing(int mode);
gdf_error gdfx_debug_occupy_sms(int blocks, unsigned microseconds);

/* Multi-GPU layer helper (libgdf_b200/dist.py): indices[i] = payload[indices[i]] in place for every
 * non-negative entry (< payload_rows), -1 otherwise.  `indices` is a GDF_INT32 join output column;
 * `payload` is a device array of the global row ids that travelled with the exchanged keys. */
gdf_error gdfx_remap_indices(gdf_column *indices, const int32_t *payload, size_t payload_rows);

/* Multi-GPU layer: split one 4/8-byte integer key column into num_partitions destination ranges of
 * {key, global row id} pairs, ids = id_base + row position.  out_keys (same width as the key) and out_ids
 * are caller-allocated device buffers of key->size elements; partition_offsets is a HOST array of
 * num_partitions + 1 start offsets.  The destination of a key is a fixed function of its value, the same on
 * every rank and independent of the hash bits the single-GPU join uses internally. */
gdf_error gdfx_partition_pairs(gdf_column *key, int32_t id_base, int num_partitions, void *out_keys,
                               int32_t *out_ids, unsigned long long *partition_offsets);

/* Multi-GPU layer: gdf_inner_join (kind 0) / gdf_left_join (kind 1) on ONE key column of exchanged rows whose
 * global row ids travel beside the keys; the GDF_INT32 output columns (library-allocated, freed with
 * gdf_column_free) hold those ids instead of positions, -1 = no partner. */
gdf_error gdfx_join_pairs(int kind, gdf_column *left_key, const int32_t *left_ids, gdf_column *right_key,
                          const int32_t *right_ids, gdf_column *out_l, gdf_column *out_r);

/* Fused partition + exchange over peer memory (multi-GPU layer, libgdf_b200/dist.py PeerExchange).
 *   gdfx_partition_count         rows of `key` per destination (host array [num_partitions])
 *   gdfx_partition_scatter_peer  one pass that writes every {key, id = id_base + position} pair straight
 *                                into its destination's receive buffers: dst_keys[p] / dst_ids[p] are device
 *                                pointers valid on THIS device - local buffers or peer buffers mapped with
 *                                gdfx_peer_open, in which case the stores cross NVLink from inside the kernel -
 *                                and dst_offsets[p] is this rank's first row inside destination p's buffers
 *                                (host arrays, num_partitions <= 16)
 *   gdfx_peer_alloc/open/close/free   cudaMalloc'd buffers shared between the ranks' processes through
 *                                64-byte CUDA IPC handles */
gdf_error gdfx_partition_count(gdf_column *key, int num_partitions, unsigned long long *counts);
gdf_error gdfx_partition_scatter_peer(gdf_column *key, int32_t id_base, int num_partitions,
                                      void * const *dst_keys, int32_t * const *dst_ids,
                                      const unsigned long long *dst_offsets);
gdf_error gdfx_peer_alloc(void **ptr, size_t bytes, char *handle64);
gdf_error gdfx_peer_open(const char *handle64, void **ptr);
gdf_error gdfx_peer_close(void *ptr);
gdf_error gdfx_peer_free(void *ptr);

/* Fused partition + exchange, second generation: ONE partition pass per side (PeerExchange.fused_inner_join).
 * The sender partitions with a COMBINED geometry - bin = destination rank * nlocal + the receiver's local radix
 * partition (nlocal a power of two, ranks * nlocal <= 256) - and stores compact {key32, id32} pairs straight into the
 * destination's pair buffer, so what a rank receives is already partition-contiguous and it joins it without a
 * histogram / scatter pass of its own.  Applies when every build key fits 32 bits (hi_or == 0 on all ranks).
 *   gdfx_xjoin_count    counts[r * nlocal + p] (host) = rows of `key` going to (rank r, local partition p);
 *                       *hi_or = OR of the keys' high words
 *   gdfx_xjoin_scatter  dst_pairs[r] = rank r's pair buffer as mapped on this device (gdfx_peer_open), 8 bytes per
 *                       pair; dst_offsets[r * nlocal + p] (host) = where this rank's pairs of that bin start; counts =
 *                       what gdfx_xjoin_count returned.  LAYOUT: every (sender, bin) slot of a receive buffer holds an
 *                       whole number of 32-byte granules (4 pairs) - the rest of a slot is padded with {0, INT_MIN} "no
 *                       row" pairs, which the local join skips - so that only sector-aligned bulk stores cross NVLink
 *                       (8-byte peer stores halve its throughput, partial sectors cost 10-30 %,
 *                       profiles/r02_p2p_store_bench.txt); offsets and per-partition totals are those of the padded
 *                       slots (dist.plan_fused_exchange / gdfx_xjoin_plan_dev do the rounding)
 *   gdfx_xjoin_local    INNER join of this rank's received pairs; *_counts[p] (host) = pairs of local partition p
 *                       (summed over the senders); outputs hold the travelling ids, as gdfx_join_pairs */
gdf_error gdfx_xjoin_count(gdf_column *key, int ranks, int nlocal, unsigned long long *counts, unsigned *hi_or);
gdf_error gdfx_xjoin_scatter(gdf_column *key, int32_t id_base, int ranks, int nlocal, void * const *dst_pairs,
                             const unsigned long long *dst_offsets, const unsigned long long *counts);
/* The same exchange WITHOUT host round trips between the histogram and the scatter: counts stay on the device
 * (d_counts[ranks * nlocal + 1]: the bins, then the OR of the keys' high words), the caller all-gathers
 * {build counts | probe counts} of every rank into d_all[ranks][2 * (ranks * nlocal + 1)], gdfx_xjoin_plan_dev turns
 * that into this rank's write offsets on the device (and raises d_status[0] = wide build keys, d_status[1] = a
 * receive buffer of cap_build / cap_probe pairs would overflow), and gdfx_xjoin_scatter_dev reads offsets and status
 * from the device - it writes nothing when a status flag is set, so the caller can check the flags AFTER launching.
 * ctas_per_sm > 0 caps the scatter's resident CTAs per SM (it is NVLink-bound; the rest of the SM is left to a kernel
 * running beside it, see gdfx_xjoin_build). */
gdf_error gdfx_xjoin_count_dev(gdf_column *key, int ranks, int nlocal, unsigned long long *d_counts);
gdf_error gdfx_xjoin_plan_dev(const unsigned long long *d_all, int ranks, int nlocal, int rank, unsigned long long cap_build,
                              unsigned long long cap_probe, unsigned long long *d_off_build, unsigned long long *d_off_probe,
                              int *d_status);
gdf_error gdfx_xjoin_scatter_dev(gdf_column *key, int32_t id_base, int ranks, int nlocal, void * const *dst_pairs,
                                 const unsigned long long *d_offsets, const unsigned long long *d_counts, const int *d_status,
                                 int ctas_per_sm);
/* gdfx_xjoin_local in two stages, so that a rank fills its hash tables WHILE the probe side is still crossing NVLink:
 * gdfx_xjoin_build is ordered after everything issued so far on the legacy stream (i.e. after the build side's
 * exchange) and, with overlap != 0, runs on a private non-blocking stream; gdfx_xjoin_probe makes the legacy stream wait
 * for it, probes, fills the outputs and releases the handle (it must be called exactly once per successful build). */
gdf_error gdfx_xjoin_build(const void *build_pairs, const unsigned long long *build_counts, int nlocal, int overlap,
                           void **handle);
gdf_error gdfx_xjoin_probe(void *handle, const void *probe_pairs, const unsigned long long *probe_counts,
                           gdf_column *out_l, gdf_column *out_r);
gdf_error gdfx_xjoin_local(const void *probe_pairs, const unsigned long long *probe_counts, const void *build_pairs,
                           const unsigned long long *build_counts, int nlocal, gdf_column *out_l, gdf_column *out_r);

/* Multi-GPU layer, composite-key joins with NULLs (C5): row validity travels with the rows as a byte
 * column.  gdfx_rows_valid_to_bytes: out[i] = 1 iff every column's validity bit i is set (no mask = all
 * valid; the reference's row-valid rule, gdf_table.cuh:63-98).  gdfx_bytes_to_valid: LSB-first Arrow bitmask
 * (ceil(rows/8) bytes) rebuilt from such a byte column on the receiving rank. */
gdf_error gdfx_rows_valid_to_bytes(gdf_column **cols, int num_cols, int8_t *out);
gdf_error gdfx_bytes_to_valid(const int8_t *in, size_t rows, gdf_valid_type *out);
