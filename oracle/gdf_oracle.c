/* TEST INFRASTRUCTURE ONLY - never linked into, imported by, or called from the product path.
 *
 * CPU restatement (plain C, scalar, single thread) of the reference algorithms on the gdf hot path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
 * The reference has NO CPU implementation (it is a CUDA library), so every function here restates
 * the semantics of the reference's kernels, citing the file:line it follows under
 * /root/reference/libgdf/src.  Pinned against: the reference's own golden vectors
 * (tests/golden/reference_vectors.json, copied from src/tests/baselines/sqls_tests_new_api.dat and
 * src/tests/cpp/sqls_tester.cu), published MurmurHash3_x86_32 known answers, and - on the GPU box -
 * the reference's own kernels compiled for sm_100a (oracle/_ref/libgdf_ref.so, see build_ref.sh).
 *
 * Conventions: columns are raw host arrays; dtype codes are gdf_dtype values (1=INT8 .. 6=FLOAT64,
 * 7=DATE32, 8=DATE64, 9=TIMESTAMP); validity masks are Arrow bitmaps, LSB first, NULL = all valid.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { T_INT8 = 1, T_INT16, T_INT32, T_INT64, T_FLOAT32, T_FLOAT64, T_DATE32, T_DATE64, T_TIMESTAMP };
enum { OP_SUM = 0, OP_MIN = 1, OP_MAX = 2, OP_AVG = 3, OP_COUNT = 4 }; /* gdf_agg_op, ref types.h:123-131 */
enum { J_INNER = 0, J_LEFT = 1, J_FULL = 2 };

static int width_of(int dtype) { /* ref column.cpp:237-275 */
  switch (dtype) {
    case T_INT8: return 1;
    case T_INT16: return 2;
    case T_INT32: case T_FLOAT32: case T_DATE32: return 4;
    case T_INT64: case T_FLOAT64: case T_DATE64: case T_TIMESTAMP: return 8;
    default: return 0;
  }
}

static int bit_is_valid(const uint8_t* mask, size_t i) { /* ref include/gdf/utils.h:10-15 */
  return mask == NULL || ((mask[i >> 3] >> (i & 7)) & 1);
}

/* ---------------------------------------------------------------------------------------------
 * MurmurHash3_x86_32, seed 0 (public-domain algorithm by Austin Appleby; the reference
 * instantiates it per column type, ref hashmap/hash_functions.cuh:31-121).
 * ------------------------------------------------------------------------------------------- */
static uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

uint32_t orc_murmur3_32(const void* key, int len) {
  const uint8_t* data = (const uint8_t*)key;
  const int nblocks = len / 4;
  uint32_t h1 = 0;
  const uint32_t c1 = 0xcc9e2d51u, c2 = 0x1b873593u;
  for (int i = 0; i < nblocks; ++i) {
    uint32_t k1;
    memcpy(&k1, data + 4 * i, 4);
    k1 *= c1; k1 = rotl32(k1, 15); k1 *= c2;
    h1 ^= k1; h1 = rotl32(h1, 13); h1 = h1 * 5 + 0xe6546b64u;
  }
  const uint8_t* tail = data + nblocks * 4;
  uint32_t k1 = 0;
  switch (len & 3) {
    case 3: k1 ^= (uint32_t)tail[2] << 16; /* fallthrough */
    case 2: k1 ^= (uint32_t)tail[1] << 8;  /* fallthrough */
    case 1: k1 ^= tail[0];
            k1 *= c1; k1 = rotl32(k1, 15); k1 *= c2; h1 ^= k1;
  }
  h1 ^= (uint32_t)len;
  h1 ^= h1 >> 16; h1 *= 0x85ebca6bu; h1 ^= h1 >> 13; h1 *= 0xc2b2ae35u; h1 ^= h1 >> 16;
  return h1;
}

/* IdentityHash: static_cast<uint32_t>(value) (ref hash_functions.cuh:156-160).  Float inputs use the
 * GPU's saturating conversion (negative/NaN -> 0, too large -> UINT32_MAX). */
static uint32_t sat_u32(double v) {
  if (!(v > 0.0)) return 0;
  if (v >= 4294967295.0) return 0xffffffffu;
  return (uint32_t)v;
}
static uint32_t identity_hash(int dtype, const void* p) {
  switch (dtype) {
    case T_INT8: return (uint32_t)(int32_t) * (const int8_t*)p;
    case T_INT16: return (uint32_t)(int32_t) * (const int16_t*)p;
    case T_INT32: case T_DATE32: return *(const uint32_t*)p;
    case T_FLOAT32: return sat_u32((double)*(const float*)p);
    case T_FLOAT64: return sat_u32(*(const double*)p);
    default: return (uint32_t) * (const uint64_t*)p;
  }
}

static uint32_t hash_combine(uint32_t lhs, uint32_t rhs) { /* ref hash_functions.cuh:66-72 */
  return lhs ^ (rhs + 0x9e3779b9u + (lhs << 6) + (lhs >> 2));
}

/* hash of one row over ncols columns (ref gdf_table.cuh:705-854): first column as is, the others
 * folded with hash_combine.  identity != 0 selects GDF_HASH_IDENTITY. */
static uint32_t row_hash(int ncols, const void* const* data, const int* dtypes, size_t row, int identity) {
  uint32_t h = 0;
  for (int c = 0; c < ncols; ++c) {
    const int w = width_of(dtypes[c]);
    const uint8_t* p = (const uint8_t*)data[c] + row * (size_t)w;
    const uint32_t hc = identity ? identity_hash(dtypes[c], p) : orc_murmur3_32(p, w);
    h = c ? hash_combine(h, hc) : hc;
  }
  return h;
}

/* gdf_hash (ref hashing.cu:83-154) */
void orc_hash_rows(int ncols, const void* const* data, const int* dtypes, size_t n, int identity, int32_t* out) {
  for (size_t r = 0; r < n; ++r) out[r] = (int32_t)row_hash(ncols, data, dtypes, r, identity);
}

/* gdf_hash_partition's row -> partition map (ref hashing.cu:196-237,259-320): hash & (n-1) for a
 * power-of-two partition count, unsigned hash % n otherwise. */
void orc_partition_ids(int ncols, const void* const* data, const int* dtypes, size_t n, int identity,
                       int num_partitions, int32_t* out) {
  const uint32_t np = (uint32_t)num_partitions;
  const int pow2 = (np & (np - 1)) == 0;
  for (size_t r = 0; r < n; ++r) {
    const uint32_t h = row_hash(ncols, data, dtypes, r, identity);
    out[r] = (int32_t)(pow2 ? (h & (np - 1)) : (h % np));
  }
}

/* typed `==` of one column value in two rows (ref gdf_table.cuh:581-691) */
static int value_equal(int dtype, const void* a, size_t ra, const void* b, size_t rb) {
  switch (dtype) {
    case T_INT8: return ((const int8_t*)a)[ra] == ((const int8_t*)b)[rb];
    case T_INT16: return ((const int16_t*)a)[ra] == ((const int16_t*)b)[rb];
    case T_INT32: case T_DATE32: return ((const int32_t*)a)[ra] == ((const int32_t*)b)[rb];
    case T_FLOAT32: return ((const float*)a)[ra] == ((const float*)b)[rb];
    case T_FLOAT64: return ((const double*)a)[ra] == ((const double*)b)[rb];
    default: return ((const int64_t*)a)[ra] == ((const int64_t*)b)[rb];
  }
}

static int row_is_valid(int ncols, const uint8_t* const* valid, size_t row) { /* ref gdf_table.cuh:63-98 */
  if (!valid) return 1;
  for (int c = 0; c < ncols; ++c)
    if (!bit_is_valid(valid[c], row)) return 0;
  return 1;
}

/* ---------------------------------------------------------------------------------------------
 * Hash join (ref join/hash/join_kernels.cuh:48-78 build, :266-455 probe; join_compute_api.h:147-186
 * full-join tail).  Build on the right table: valid rows are chained by row hash; each valid left
 * row walks its chain comparing hash then rows_equal and emits (l, r) per match; LEFT/FULL emit
 * (l,-1) when nothing matched (NULL-key rows never match); FULL appends (-1, r) for every right row
 * that never appeared in the right output.  Pair ORDER is unspecified in the reference (global atomic
 * cursors) - compare results as multisets.  Returns the number of pairs, or -1 if `capacity` is too
 * small (call again with a larger buffer).
 * ------------------------------------------------------------------------------------------- */
long long orc_join(int kind, int ncols, const int* dtypes, const void* const* ldata, const uint8_t* const* lvalid,
                   size_t nl, const void* const* rdata, const uint8_t* const* rvalid, size_t nr,
                   int32_t* out_l, int32_t* out_r, size_t capacity) {
  size_t nbuckets = 1;
  while (nbuckets < 2 * nr + 1) nbuckets <<= 1; /* ref join_compute_api.h:386: 50 % occupancy */
  int32_t* head = (int32_t*)malloc(nbuckets * sizeof(int32_t));
  int32_t* next = (int32_t*)malloc((nr ? nr : 1) * sizeof(int32_t));
  uint32_t* rh = (uint32_t*)malloc((nr ? nr : 1) * sizeof(uint32_t));
  uint8_t* matched = (uint8_t*)calloc(nr ? nr : 1, 1);
  if (!head || !next || !rh || !matched) { free(head); free(next); free(rh); free(matched); return -2; }
  for (size_t i = 0; i < nbuckets; ++i) head[i] = -1;
  for (size_t r = nr; r-- > 0;) { /* reverse so chains list rows in ascending order */
    if (!row_is_valid(ncols, rvalid, r)) continue;
    rh[r] = row_hash(ncols, rdata, dtypes, r, 0);
    const size_t b = rh[r] & (nbuckets - 1);
    next[r] = head[b];
    head[b] = (int32_t)r;
  }
  long long count = 0;
  int overflow = 0;
  for (size_t l = 0; l < nl; ++l) {
    int found = 0;
    if (row_is_valid(ncols, lvalid, l)) {
      const uint32_t h = row_hash(ncols, ldata, dtypes, l, 0);
      for (int32_t r = head[h & (nbuckets - 1)]; r >= 0; r = next[r]) {
        if (rh[r] != h) continue;
        int eq = 1;
        for (int c = 0; c < ncols && eq; ++c) eq = value_equal(dtypes[c], ldata[c], l, rdata[c], (size_t)r);
        if (!eq) continue;
        found = 1;
        matched[r] = 1;
        if ((size_t)count < capacity) { out_l[count] = (int32_t)l; out_r[count] = r; } else overflow = 1;
        ++count;
      }
    }
    if (!found && kind != J_INNER) {
      if ((size_t)count < capacity) { out_l[count] = (int32_t)l; out_r[count] = -1; } else overflow = 1;
      ++count;
    }
  }
  if (kind == J_FULL) {
    for (size_t r = 0; r < nr; ++r) {
      if (matched[r]) continue;
      if ((size_t)count < capacity) { out_l[count] = -1; out_r[count] = (int32_t)r; } else overflow = 1;
      ++count;
    }
  }
  free(head); free(next); free(rh); free(matched);
  return overflow ? -1 : count;
}

/* ---------------------------------------------------------------------------------------------
 * Hash group-by with one aggregation column (ref groupby/hash/groupby_kernels.cuh:47-160,
 * groupby/hash/aggregation_operations.cuh:30-74, groupby/groupby.cuh:88-190,308-386).
 * Groups are emitted in first-appearance order here; the reference's order is unspecified.
 *  - SUM/MIN/MAX aggregate in the INPUT column's C type (int8 sums wrap) and the output buffer is
 *    written as that type;
 *  - COUNT counts in the OUTPUT column's C type (out_dtype);
 *  - AVG = sum (input type, wrapped) / (avg_type)count, converted to the output type.
 * Returns the number of groups.
 * ------------------------------------------------------------------------------------------- */
typedef union { int64_t i; double d; float f; } acc_t;

static double load_as_double(int dtype, const void* p, size_t r) {
  switch (dtype) {
    case T_INT8: return ((const int8_t*)p)[r];
    case T_INT16: return ((const int16_t*)p)[r];
    case T_INT32: case T_DATE32: return ((const int32_t*)p)[r];
    case T_FLOAT32: return ((const float*)p)[r];
    case T_FLOAT64: return ((const double*)p)[r];
    default: return (double)((const int64_t*)p)[r];
  }
}
static int64_t load_as_i64(int dtype, const void* p, size_t r) {
  switch (dtype) {
    case T_INT8: return ((const int8_t*)p)[r];
    case T_INT16: return ((const int16_t*)p)[r];
    case T_INT32: case T_DATE32: return ((const int32_t*)p)[r];
    default: return ((const int64_t*)p)[r];
  }
}
static int is_float(int dtype) { return dtype == T_FLOAT32 || dtype == T_FLOAT64; }

static int64_t wrap_int(int dtype, int64_t v) { /* arithmetic in the column's own width */
  switch (width_of(dtype)) {
    case 1: return (int8_t)v;
    case 2: return (int16_t)v;
    case 4: return (int32_t)v;
    default: return v;
  }
}

static void store_typed(int dtype, void* out, size_t at, int64_t iv, double dv, int from_float) {
  switch (dtype) {
    case T_INT8: ((int8_t*)out)[at] = from_float ? (int8_t)dv : (int8_t)iv; break;
    case T_INT16: ((int16_t*)out)[at] = from_float ? (int16_t)dv : (int16_t)iv; break;
    case T_INT32: case T_DATE32: ((int32_t*)out)[at] = from_float ? (int32_t)dv : (int32_t)iv; break;
    case T_FLOAT32: ((float*)out)[at] = from_float ? (float)dv : (float)iv; break;
    case T_FLOAT64: ((double*)out)[at] = from_float ? dv : (double)iv; break;
    default: ((int64_t*)out)[at] = from_float ? (int64_t)dv : iv; break;
  }
}

long long orc_groupby(int op, int ncols, const int* dtypes, const void* const* keys, size_t n,
                      int val_dtype, const void* values, int out_dtype, void* const* out_keys, void* out_agg) {
  size_t nbuckets = 1;
  while (nbuckets < 2 * n + 1) nbuckets <<= 1;
  int32_t* slot_group = (int32_t*)malloc(nbuckets * sizeof(int32_t));
  int32_t* first_row = (int32_t*)malloc((n ? n : 1) * sizeof(int32_t));
  acc_t* acc = (acc_t*)malloc((n ? n : 1) * sizeof(acc_t));
  int64_t* cnt = (int64_t*)malloc((n ? n : 1) * sizeof(int64_t));
  if (!slot_group || !first_row || !acc || !cnt) { free(slot_group); free(first_row); free(acc); free(cnt); return -2; }
  for (size_t i = 0; i < nbuckets; ++i) slot_group[i] = -1;
  const int acc_dtype = (op == OP_COUNT) ? out_dtype : val_dtype;
  const int fl = is_float(acc_dtype);
  const int f32 = acc_dtype == T_FLOAT32;
  long long ngroups = 0;
  for (size_t r = 0; r < n; ++r) {
    const uint32_t h = row_hash(ncols, keys, dtypes, r, 0);
    size_t s = h & (nbuckets - 1);
    int32_t g;
    while (1) { /* linear probing, key = first row of the group (ref concurrent_unordered_map.cuh:485-544) */
      g = slot_group[s];
      if (g < 0) {
        g = (int32_t)ngroups++;
        slot_group[s] = g;
        first_row[g] = (int32_t)r;
        cnt[g] = 0;
        /* identities (ref aggregation_operations.cuh:32,44,55,66) */
        if (fl) {
          const double lo = f32 ? -3.402823466e+38 : -1.7976931348623157e+308;
          acc[g].d = op == OP_MIN ? -lo : (op == OP_MAX ? lo : 0.0);
        } else {
          int64_t mx, mn;
          switch (width_of(acc_dtype)) {
            case 1: mx = INT8_MAX; mn = INT8_MIN; break;
            case 2: mx = INT16_MAX; mn = INT16_MIN; break;
            case 4: mx = INT32_MAX; mn = INT32_MIN; break;
            default: mx = INT64_MAX; mn = INT64_MIN; break;
          }
          acc[g].i = op == OP_MIN ? mx : (op == OP_MAX ? mn : 0);
        }
        break;
      }
      int eq = 1;
      for (int c = 0; c < ncols && eq; ++c) eq = value_equal(dtypes[c], keys[c], r, keys[c], (size_t)first_row[g]);
      if (eq) break;
      s = (s + 1) & (nbuckets - 1);
    }
    cnt[g] += 1;
    if (op == OP_COUNT) {
      if (fl) acc[g].d = f32 ? (double)(float)((float)acc[g].d + 1.0f) : acc[g].d + 1.0;
      else acc[g].i = wrap_int(acc_dtype, acc[g].i + 1);
    } else if (fl) {
      const double v = load_as_double(val_dtype, values, r);
      if (op == OP_MIN) acc[g].d = v < acc[g].d ? v : acc[g].d;
      else if (op == OP_MAX) acc[g].d = v > acc[g].d ? v : acc[g].d;
      else acc[g].d = f32 ? (double)(float)((float)acc[g].d + (float)v) : acc[g].d + v;
    } else {
      const int64_t v = load_as_i64(val_dtype, values, r);
      if (op == OP_MIN) acc[g].i = v < acc[g].i ? v : acc[g].i;
      else if (op == OP_MAX) acc[g].i = v > acc[g].i ? v : acc[g].i;
      else acc[g].i = wrap_int(acc_dtype, (int64_t)((uint64_t)acc[g].i + (uint64_t)v));
    }
  }
  for (long long g = 0; g < ngroups; ++g) {
    for (int c = 0; c < ncols; ++c) { /* copy_row (ref gdf_table.cuh:474-567) */
      const int w = width_of(dtypes[c]);
      memcpy((uint8_t*)out_keys[c] + (size_t)g * w, (const uint8_t*)keys[c] + (size_t)first_row[g] * w, (size_t)w);
    }
    if (op == OP_AVG) {
      /* ref groupby.cuh:308-328: avg = sum / static_cast<avg_type>(count), evaluated under the usual
       * arithmetic conversions of (sum_type, avg_type) and then converted to avg_type. */
      const int any_f64 = val_dtype == T_FLOAT64 || out_dtype == T_FLOAT64;
      const int any_f32 = val_dtype == T_FLOAT32 || out_dtype == T_FLOAT32;
      if (any_f64 || any_f32) {
        double sum = fl ? acc[g].d : (double)acc[g].i;
        double c = is_float(out_dtype) ? (out_dtype == T_FLOAT32 ? (double)(float)cnt[g] : (double)cnt[g])
                                       : (double)wrap_int(out_dtype, cnt[g]);
        double q;
        if (any_f64) q = sum / c;
        else q = (double)((fl ? (float)acc[g].d : (float)acc[g].i) / (float)c); /* float arithmetic */
        store_typed(out_dtype, out_agg, (size_t)g, 0, q, 1);
      } else { /* integer / integer (int or int64 arithmetic give the same quotient) */
        const int64_t c = wrap_int(out_dtype, cnt[g]);
        const int64_t q = c ? acc[g].i / c : 0;
        store_typed(out_dtype, out_agg, (size_t)g, q, 0.0, 0);
      }
    } else {
      store_typed(acc_dtype, out_agg, (size_t)g, fl ? 0 : acc[g].i, fl ? acc[g].d : 0.0, fl);
    }
  }
  free(slot_group); free(first_row); free(acc); free(cnt);
  return ngroups;
}

/* ---------------------------------------------------------------------------------------------
 * gdf_filter on int64 columns (ref sqls_rtti_comp.hpp:200-213,343-370): ascending indices of rows
 * where no column differs from its comparand.  Typed variant used for the bench's CPU baseline;
 * the general (mixed dtype) restatement lives in oracle/np_oracle.py.
 * ------------------------------------------------------------------------------------------- */
size_t orc_filter_i64(const int64_t* data, size_t n, int64_t value, uint64_t* out_idx) {
  size_t k = 0;
  for (size_t i = 0; i < n; ++i)
    if (!(data[i] != value)) out_idx[k++] = i;
  return k;
}

/* gdf_sum_i64 with validity (ref reductions.cu:26-61): NULL -> identity, wrapping int64 sum */
int64_t orc_sum_i64(const int64_t* data, const uint8_t* valid, size_t n) {
  uint64_t s = 0;
  for (size_t i = 0; i < n; ++i)
    if (bit_is_valid(valid, i)) s += (uint64_t)data[i];
  return (int64_t)s;
}
