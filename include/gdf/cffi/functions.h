/* Hot-path entry points of the gdf_* C ABI (drop-in boundary).  Plain C, include-free, cdef()'d verbatim by
 * libgdf_b200/libgdf_cffi.  Each prototype is character-for-character ABI compatible with the reference
 * declaration cited beside it (reference: libgdf/include/gdf/cffi/functions.h); the definitions are
 * from-scratch sm_100a CUDA in libgdf_b200/csrc/.  Functions of the reference that are outside the
 * hot path (unary math, casts, datetime, sort plans, IPC, CSV, quantiles ...) are
 * intentionally not declared and not exported.
 */

/* ---- profiling ranges (ref functions.h:18-52, src/nvtx_utils.cpp:19-71) ---- */
gdf_error gdf_nvtx_range_push(char const * const name, gdf_color color);
gdf_error gdf_nvtx_range_push_hex(char const * const name, unsigned int color);
gdf_error gdf_nvtx_range_pop();

/* ---- column / context / error plumbing (ref functions.h:54-107, src/column.cpp:160-275,
 *      src/context.cpp:3-11, src/errorhandling.cpp:5-35, src/cudautils.cu:4-14) ---- */
gdf_error gdf_count_nonzero_mask(gdf_valid_type const * masks, int num_rows, int * count);
gdf_size_type gdf_column_sizeof();
gdf_error gdf_column_view(gdf_column *column, void *data, gdf_valid_type *valid,
                          gdf_size_type size, gdf_dtype dtype);
gdf_error gdf_column_view_augmented(gdf_column *column, void *data, gdf_valid_type *valid,
                                    gdf_size_type size, gdf_dtype dtype, gdf_size_type null_count);
gdf_error gdf_column_free(gdf_column *column);
gdf_error gdf_context_view(gdf_context *context, int flag_sorted, gdf_method flag_method,
                           int flag_distinct, int flag_sort_result, int flag_sort_inplace);
const char * gdf_error_get_name(gdf_error errcode);
int gdf_cuda_last_error();
const char * gdf_cuda_error_string(int cuda_error);
const char * gdf_cuda_error_name(int cuda_error);
gdf_error get_column_byte_width(gdf_column * col, int * width);

/* ---- concatenation (ref functions.h:94, src/column.cpp:53-153, src/validops.cu:203-256): what a
 * caller uses to stitch per-GPU shards of a column back together.  output->size must equal the sum of the input
 * sizes and its buffers are caller-allocated; a column without mask counts as all valid.  gdf_mask_concat has no
 * public prototype in the reference (column.cpp:29 declares it locally); masks_to_concat / column_lengths may be
 * host, managed or device arrays. ---- */
gdf_error gdf_column_concat(gdf_column *output, gdf_column *columns_to_concat[], int num_columns);
gdf_error gdf_mask_concat(gdf_valid_type *output_mask, gdf_size_type output_column_length,
                          gdf_valid_type *masks_to_concat[], gdf_size_type *column_lengths,
                          gdf_size_type num_columns);

/* ---- hash joins (ref functions.h:226-318, src/join/joining.cu:571-653).
 * Output index columns are allocated by the library with rmmAlloc and released by the caller with
 * gdf_column_free; unmatched side is -1; pair order is unspecified. ---- */
gdf_error gdf_inner_join(gdf_column **left_cols, int num_left_cols, int left_join_cols[],
                         gdf_column **right_cols, int num_right_cols, int right_join_cols[],
                         int num_cols_to_join, int result_num_cols, gdf_column **result_cols,
                         gdf_column * left_indices, gdf_column * right_indices,
                         gdf_context *join_context);
gdf_error gdf_left_join(gdf_column **left_cols, int num_left_cols, int left_join_cols[],
                         gdf_column **right_cols, int num_right_cols, int right_join_cols[],
                         int num_cols_to_join, int result_num_cols, gdf_column **result_cols,
                         gdf_column * left_indices, gdf_column * right_indices,
                         gdf_context *join_context);
gdf_error gdf_full_join(gdf_column **left_cols, int num_left_cols, int left_join_cols[],
                         gdf_column **right_cols, int num_right_cols, int right_join_cols[],
                         int num_cols_to_join, int result_num_cols, gdf_column **result_cols,
                         gdf_column * left_indices, gdf_column * right_indices,
                         gdf_context *join_context);

/* ---- row hash + hash partition (ref functions.h:344-351,378; src/hashing.cu:83-154,559-654) ---- */
gdf_error gdf_hash_partition(int num_input_cols, gdf_column * input[], int columns_to_hash[],
                             int num_cols_to_hash, int num_partitions,
                             gdf_column * partitioned_output[], int partition_offsets[],
                             gdf_hash_func hash);
gdf_error gdf_hash(int num_cols, gdf_column **input, gdf_hash_func hash, gdf_column *output);

/* ---- element-wise binary ops (ref functions.h:528-624, src/binaryops.cu:168-526) ---- */
gdf_error gdf_add_generic(gdf_column *lhs, gdf_column *rhs, gdf_column *output);
gdf_error gdf_add_i32(gdf_column *lhs, gdf_column *rhs, gdf_column *output);
gdf_error gdf_add_i64(gdf_column *lhs, gdf_column *rhs, gdf_column *output);
gdf_error gdf_add_f32(gdf_column *lhs, gdf_column *rhs, gdf_column *output);
gdf_error gdf_add_f64(gdf_column *lhs, gdf_column *rhs, gdf_column *output);
gdf_error gdf_sub_generic(gdf_column *lhs, gdf_column *rhs, gdf_column *output);
gdf_error gdf_sub_i32(gdf_column *lhs, gdf_column *rhs, gdf_column *output);
gdf_error gdf_sub_i64(gdf_column *lhs, gdf_column *rhs, gdf_column *output);
gdf_error gdf_sub_f32(gdf_column *lhs, gdf_column *rhs, gdf_column *output);
gdf_error gdf_sub_f64(gdf_column *lhs, gdf_column *rhs, gdf_column *output);
gdf_error gdf_mul_generic(gdf_column *lhs, gdf_column *rhs, gdf_column *output);
gdf_error gdf_mul_i32(gdf_column *lhs, gdf_column *rhs, gdf_column *output);
gdf_error gdf_mul_i64(gdf_column *lhs, gdf_column *rhs, gdf_column *output);
gdf_error gdf_mul_f32(gdf_column *lhs, gdf_column *rhs, gdf_column *output);
gdf_error gdf_mul_f64(gdf_column *lhs, gdf_column *rhs, gdf_column *output);
gdf_error gdf_floordiv_generic(gdf_column *lhs, gdf_column *rhs, gdf_column *output);
gdf_error gdf_floordiv_i32(gdf_column *lhs, gdf_column *rhs, gdf_column *output);
gdf_error gdf_floordiv_i64(gdf_column *lhs, gdf_column *rhs, gdf_column *output);
gdf_error gdf_floordiv_f32(gdf_column *lhs, gdf_column *rhs, gdf_column *output);
gdf_error gdf_floordiv_f64(gdf_column *lhs, gdf_column *rhs, gdf_column *output);
gdf_error gdf_div_generic(gdf_column *lhs, gdf_column *rhs, gdf_column *output);
gdf_error gdf_div_f32(gdf_column *lhs, gdf_column *rhs, gdf_column *output);
gdf_error gdf_div_f64(gdf_column *lhs, gdf_column *rhs, gdf_column *output);
gdf_error gdf_gt_generic(gdf_column *lhs, gdf_column *rhs, gdf_column *output);
gdf_error gdf_gt_i8(gdf_column *lhs, gdf_column *rhs, gdf_column *output);
gdf_error gdf_gt_i32(gdf_column *lhs, gdf_column *rhs, gdf_column *output);
gdf_error gdf_gt_i64(gdf_column *lhs, gdf_column *rhs, gdf_column *output);
gdf_error gdf_gt_f32(gdf_column *lhs, gdf_column *rhs, gdf_column *output);
gdf_error gdf_gt_f64(gdf_column *lhs, gdf_column *rhs, gdf_column *output);
gdf_error gdf_ge_generic(gdf_column *lhs, gdf_column *rhs, gdf_column *output);
gdf_error gdf_ge_i8(gdf_column *lhs, gdf_column *rhs, gdf_column *output);
gdf_error gdf_ge_i32(gdf_column *lhs, gdf_column *rhs, gdf_column *output);
gdf_error gdf_ge_i64(gdf_column *lhs, gdf_column *rhs, gdf_column *output);
gdf_error gdf_ge_f32(gdf_column *lhs, gdf_column *rhs, gdf_column *output);
gdf_error gdf_ge_f64(gdf_column *lhs, gdf_column *rhs, gdf_column *output);
gdf_error gdf_lt_generic(gdf_column *lhs, gdf_column *rhs, gdf_column *output);
gdf_error gdf_lt_i8(gdf_column *lhs, gdf_column *rhs, gdf_column *output);
gdf_error gdf_lt_i32(gdf_column *lhs, gdf_column *rhs, gdf_column *output);
gdf_error gdf_lt_i64(gdf_column *lhs, gdf_column *rhs, gdf_column *output);
gdf_error gdf_lt_f32(gdf_column *lhs, gdf_column *rhs, gdf_column *output);
gdf_error gdf_lt_f64(gdf_column *lhs, gdf_column *rhs, gdf_column *output);
gdf_error gdf_le_generic(gdf_column *lhs, gdf_column *rhs, gdf_column *output);
gdf_error gdf_le_i8(gdf_column *lhs, gdf_column *rhs, gdf_column *output);
gdf_error gdf_le_i32(gdf_column *lhs, gdf_column *rhs, gdf_column *output);
gdf_error gdf_le_i64(gdf_column *lhs, gdf_column *rhs, gdf_column *output);
gdf_error gdf_le_f32(gdf_column *lhs, gdf_column *rhs, gdf_column *output);
gdf_error gdf_le_f64(gdf_column *lhs, gdf_column *rhs, gdf_column *output);
gdf_error gdf_eq_generic(gdf_column *lhs, gdf_column *rhs, gdf_column *output);
gdf_error gdf_eq_i8(gdf_column *lhs, gdf_column *rhs, gdf_column *output);
gdf_error gdf_eq_i32(gdf_column *lhs, gdf_column *rhs, gdf_column *output);
gdf_error gdf_eq_i64(gdf_column *lhs, gdf_column *rhs, gdf_column *output);
gdf_error gdf_eq_f32(gdf_column *lhs, gdf_column *rhs, gdf_column *output);
gdf_error gdf_eq_f64(gdf_column *lhs, gdf_column *rhs, gdf_column *output);
gdf_error gdf_ne_generic(gdf_column *lhs, gdf_column *rhs, gdf_column *output);
gdf_error gdf_ne_i8(gdf_column *lhs, gdf_column *rhs, gdf_column *output);
gdf_error gdf_ne_i32(gdf_column *lhs, gdf_column *rhs, gdf_column *output);
gdf_error gdf_ne_i64(gdf_column *lhs, gdf_column *rhs, gdf_column *output);
gdf_error gdf_ne_f32(gdf_column *lhs, gdf_column *rhs, gdf_column *output);
gdf_error gdf_ne_f64(gdf_column *lhs, gdf_column *rhs, gdf_column *output);
gdf_error gdf_bitwise_and_generic(gdf_column *lhs, gdf_column *rhs, gdf_column *output);
gdf_error gdf_bitwise_and_i8(gdf_column *lhs, gdf_column *rhs, gdf_column *output);
gdf_error gdf_bitwise_and_i32(gdf_column *lhs, gdf_column *rhs, gdf_column *output);
gdf_error gdf_bitwise_and_i64(gdf_column *lhs, gdf_column *rhs, gdf_column *output);
gdf_error gdf_bitwise_or_generic(gdf_column *lhs, gdf_column *rhs, gdf_column *output);
gdf_error gdf_bitwise_or_i8(gdf_column *lhs, gdf_column *rhs, gdf_column *output);
gdf_error gdf_bitwise_or_i32(gdf_column *lhs, gdf_column *rhs, gdf_column *output);
gdf_error gdf_bitwise_or_i64(gdf_column *lhs, gdf_column *rhs, gdf_column *output);
gdf_error gdf_bitwise_xor_generic(gdf_column *lhs, gdf_column *rhs, gdf_column *output);
gdf_error gdf_bitwise_xor_i8(gdf_column *lhs, gdf_column *rhs, gdf_column *output);
gdf_error gdf_bitwise_xor_i32(gdf_column *lhs, gdf_column *rhs, gdf_column *output);
gdf_error gdf_bitwise_xor_i64(gdf_column *lhs, gdf_column *rhs, gdf_column *output);
gdf_error gdf_validity_and(gdf_column *lhs, gdf_column *rhs, gdf_column *output);

/* ---- columnar reductions (ref functions.h:626-666, src/reductions.cu:224-269).
 * dev_result is device memory of dev_result_size elements; the answer lands in dev_result[0]. ---- */
unsigned int gdf_reduce_optimal_output_size();
gdf_error gdf_sum_generic(gdf_column *col, void *dev_result, gdf_size_type dev_result_size);
gdf_error gdf_sum_f64(gdf_column *col, double *dev_result, gdf_size_type dev_result_size);
gdf_error gdf_sum_f32(gdf_column *col, float *dev_result, gdf_size_type dev_result_size);
gdf_error gdf_sum_i64(gdf_column *col, int64_t *dev_result, gdf_size_type dev_result_size);
gdf_error gdf_sum_i32(gdf_column *col, int32_t *dev_result, gdf_size_type dev_result_size);
gdf_error gdf_sum_i8(gdf_column *col, int8_t *dev_result, gdf_size_type dev_result_size);
gdf_error gdf_product_generic(gdf_column *col, void *dev_result, gdf_size_type dev_result_size);
gdf_error gdf_product_f64(gdf_column *col, double *dev_result, gdf_size_type dev_result_size);
gdf_error gdf_product_f32(gdf_column *col, float *dev_result, gdf_size_type dev_result_size);
gdf_error gdf_product_i64(gdf_column *col, int64_t *dev_result, gdf_size_type dev_result_size);
gdf_error gdf_product_i32(gdf_column *col, int32_t *dev_result, gdf_size_type dev_result_size);
gdf_error gdf_product_i8(gdf_column *col, int8_t *dev_result, gdf_size_type dev_result_size);
gdf_error gdf_sum_squared_generic(gdf_column *col, void *dev_result, gdf_size_type dev_result_size);
gdf_error gdf_sum_squared_f64(gdf_column *col, double *dev_result, gdf_size_type dev_result_size);
gdf_error gdf_sum_squared_f32(gdf_column *col, float *dev_result, gdf_size_type dev_result_size);
gdf_error gdf_min_generic(gdf_column *col, void *dev_result, gdf_size_type dev_result_size);
gdf_error gdf_min_f64(gdf_column *col, double *dev_result, gdf_size_type dev_result_size);
gdf_error gdf_min_f32(gdf_column *col, float *dev_result, gdf_size_type dev_result_size);
gdf_error gdf_min_i64(gdf_column *col, int64_t *dev_result, gdf_size_type dev_result_size);
gdf_error gdf_min_i32(gdf_column *col, int32_t *dev_result, gdf_size_type dev_result_size);
gdf_error gdf_min_i8(gdf_column *col, int8_t *dev_result, gdf_size_type dev_result_size);
gdf_error gdf_max_generic(gdf_column *col, void *dev_result, gdf_size_type dev_result_size);
gdf_error gdf_max_f64(gdf_column *col, double *dev_result, gdf_size_type dev_result_size);
gdf_error gdf_max_f32(gdf_column *col, float *dev_result, gdf_size_type dev_result_size);
gdf_error gdf_max_i64(gdf_column *col, int64_t *dev_result, gdf_size_type dev_result_size);
gdf_error gdf_max_i32(gdf_column *col, int32_t *dev_result, gdf_size_type dev_result_size);
gdf_error gdf_max_i8(gdf_column *col, int8_t *dev_result, gdf_size_type dev_result_size);

/* ---- predicate -> int8 stencil, stream compaction (ref functions.h:677-690,
 *      src/filterops.cu:162-662, src/streamcompactionops.cu:208-339) ---- */
gdf_error gpu_comparison_static_i8(gdf_column *lhs, int8_t value, gdf_column *output, gdf_comparison_operator operation);
gdf_error gpu_comparison_static_i16(gdf_column *lhs, int16_t value, gdf_column *output, gdf_comparison_operator operation);
gdf_error gpu_comparison_static_i32(gdf_column *lhs, int32_t value, gdf_column *output, gdf_comparison_operator operation);
gdf_error gpu_comparison_static_i64(gdf_column *lhs, int64_t value, gdf_column *output, gdf_comparison_operator operation);
gdf_error gpu_comparison_static_f32(gdf_column *lhs, float value, gdf_column *output, gdf_comparison_operator operation);
gdf_error gpu_comparison_static_f64(gdf_column *lhs, double value, gdf_column *output, gdf_comparison_operator operation);
gdf_error gpu_comparison(gdf_column *lhs, gdf_column *rhs, gdf_column *output, gdf_comparison_operator operation);
gdf_error gpu_apply_stencil(gdf_column *lhs, gdf_column * stencil, gdf_column * output);

/* ---- multi-column ORDER BY (ref functions.h:711-716, src/sqls_ops.cu:1373-1392).
 * cols is a HOST array of gdf_column structs (no masks); d_cols/d_types are caller device scratch that the call
 * fills; d_indx receives the row indices (size_t) in lexicographic ascending order of the columns. ---- */
gdf_error gdf_order_by(size_t nrows, gdf_column* cols, size_t ncols, void** d_cols, int* d_types, size_t* d_indx);

/* ---- multi-column WHERE (ref functions.h:718-725, src/sqls_ops.cu:1401-1424).
 * cols is a HOST array of gdf_column structs; d_cols/d_types are caller device scratch that the call
 * fills; d_vals is a device array of device pointers to one comparand per column; d_indx receives the
 * ascending row indices (size_t); *new_sz (host) the count. ---- */
gdf_error gdf_filter(size_t nrows, gdf_column* cols, size_t ncols, void** d_cols, int* d_types,
                     void** d_vals, size_t* d_indx, size_t* new_sz);

/* ---- hash group-by with one aggregation column (ref functions.h:727-772,
 *      src/sqls_ops.cu:1085-1487, src/groupby/groupby.cuh:210-419).  Outputs are caller-preallocated
 * at input size; the call writes the group keys / aggregates and sets their size to the group count. ---- */
gdf_error gdf_group_by_sum(int ncols, gdf_column** cols, gdf_column* col_agg,
                           gdf_column* out_col_indices, gdf_column** out_col_values,
                           gdf_column* out_col_agg, gdf_context* ctxt);
gdf_error gdf_group_by_min(int ncols, gdf_column** cols, gdf_column* col_agg,
                           gdf_column* out_col_indices, gdf_column** out_col_values,
                           gdf_column* out_col_agg, gdf_context* ctxt);
gdf_error gdf_group_by_max(int ncols, gdf_column** cols, gdf_column* col_agg,
                           gdf_column* out_col_indices, gdf_column** out_col_values,
                           gdf_column* out_col_agg, gdf_context* ctxt);
gdf_error gdf_group_by_avg(int ncols, gdf_column** cols, gdf_column* col_agg,
                           gdf_column* out_col_indices, gdf_column** out_col_values,
                           gdf_column* out_col_agg, gdf_context* ctxt);
gdf_error gdf_group_by_count(int ncols, gdf_column** cols, gdf_column* col_agg,
                           gdf_column* out_col_indices, gdf_column** out_col_values,
                           gdf_column* out_col_agg, gdf_context* ctxt);
