/* ABI types of the gdf_* hot path.  Plain C, include-free: this file is cdef()'d verbatim by the cffi
 * binding (libgdf_b200/libgdf_cffi), exactly as the reference binds its own header
 * (reference: libgdf/python/libgdf_cffi/libgdf_build.py:4-8).
 *
 * Every enum value, field order and field width below is part of the binary contract with
 * reference libgdf/include/gdf/cffi/types.h (cited per item); only the layout of this text is ours.
 * sizeof(gdf_column) == 56 (data@0 valid@8 size@16 dtype@24 null_count@32 dtype_info@40 col_name@48).
 */
typedef size_t gdf_size_type;            /* ref types.h:3 */
typedef gdf_size_type gdf_index_type;    /* ref types.h:4 */
typedef unsigned char gdf_valid_type;    /* ref types.h:5 - 1 bit per row, LSB first (ref utils.h:10-15) */
typedef long gdf_date64;
typedef int gdf_date32;
typedef int gdf_category;

/* ref types.h:15-29 */
typedef enum {
  GDF_invalid = 0, GDF_INT8, GDF_INT16, GDF_INT32, GDF_INT64, GDF_FLOAT32, GDF_FLOAT64,
  GDF_DATE32, GDF_DATE64, GDF_TIMESTAMP, GDF_CATEGORY, GDF_STRING, N_GDF_TYPES
} gdf_dtype;

/* ref types.h:39-64 - keep in step with gdf_error_get_name */
typedef enum {
  GDF_SUCCESS = 0,
  GDF_CUDA_ERROR,
  GDF_UNSUPPORTED_DTYPE,
  GDF_COLUMN_SIZE_MISMATCH,
  GDF_COLUMN_SIZE_TOO_BIG,
  GDF_DATASET_EMPTY,
  GDF_VALIDITY_MISSING,
  GDF_VALIDITY_UNSUPPORTED,
  GDF_INVALID_API_CALL,
  GDF_JOIN_DTYPE_MISMATCH,
  GDF_JOIN_TOO_MANY_COLUMNS,
  GDF_DTYPE_MISMATCH,
  GDF_UNSUPPORTED_METHOD,
  GDF_INVALID_AGGREGATOR,
  GDF_INVALID_HASH_FUNCTION,
  GDF_PARTITION_DTYPE_MISMATCH,
  GDF_HASH_TABLE_INSERT_FAILURE,
  GDF_UNSUPPORTED_JOIN_TYPE,
  GDF_C_ERROR,
  GDF_FILE_ERROR,
  GDF_MEMORYMANAGER_ERROR,
  GDF_UNDEFINED_NVTX_COLOR,
  GDF_NULL_NVTX_NAME,
  N_GDF_ERRORS
} gdf_error;

/* ref types.h:66-69 */
typedef enum { GDF_HASH_MURMUR3 = 0, GDF_HASH_IDENTITY } gdf_hash_func;

/* ref types.h:71-77 */
typedef enum { TIME_UNIT_NONE = 0, TIME_UNIT_s, TIME_UNIT_ms, TIME_UNIT_us, TIME_UNIT_ns } gdf_time_unit;

/* ref types.h:79-82 */
typedef struct { gdf_time_unit time_unit; } gdf_dtype_extra_info;

/* ref types.h:84-92 - the column descriptor: host struct, device buffers */
typedef struct gdf_column_ {
  void *data;
  gdf_valid_type *valid;
  gdf_size_type size;
  gdf_dtype dtype;
  gdf_size_type null_count;
  gdf_dtype_extra_info dtype_info;
  char *col_name;
} gdf_column;

/* ref types.h:101-105 */
typedef enum { GDF_SORT = 0, GDF_HASH, N_GDF_METHODS } gdf_method;

/* ref types.h:107-114 */
typedef enum {
  GDF_QUANT_LINEAR = 0, GDF_QUANT_LOWER, GDF_QUANT_HIGHER, GDF_QUANT_MIDPOINT, GDF_QUANT_NEAREST,
  N_GDF_QUANT_METHODS
} gdf_quantile_method;

/* ref types.h:123-131 */
typedef enum {
  GDF_SUM = 0, GDF_MIN, GDF_MAX, GDF_AVG, GDF_COUNT, GDF_COUNT_DISTINCT, N_GDF_AGG_OPS
} gdf_agg_op;

/* ref types.h:142-153 */
typedef enum {
  GDF_GREEN = 0, GDF_BLUE, GDF_YELLOW, GDF_PURPLE, GDF_CYAN, GDF_RED, GDF_WHITE, GDF_DARK_GREEN,
  GDF_ORANGE, GDF_NUM_COLORS
} gdf_color;

/* ref types.h:161-167 */
typedef struct gdf_context_ {
  int flag_sorted;
  gdf_method flag_method;
  int flag_distinct;
  int flag_sort_result;
  int flag_sort_inplace;
} gdf_context;

/* ref types.h:183-186 */
typedef enum { GDF_ORDER_ASC, GDF_ORDER_DESC } order_by_type;

/* ref types.h:188-195 */
typedef enum {
  GDF_EQUALS, GDF_NOT_EQUALS, GDF_LESS_THAN, GDF_LESS_THAN_OR_EQUALS, GDF_GREATER_THAN,
  GDF_GREATER_THAN_OR_EQUALS
} gdf_comparison_operator;
