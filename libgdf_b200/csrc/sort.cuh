// Row-order utilities shared by gdf_order_by, the sort-based group-by and the hash group-by's flag_sort_result
// (sort.cu).  Everything runs on the legacy default stream and performs no host synchronisation.
#pragma once
#include "common.cuh"

namespace b200 {

// d_perm[0..n) = the row ids 0..n-1 ordered lexicographically (ascending, column 0 most significant) by the typed
// values of `cols`; ties keep ascending row order (stable).  Integers / dates / timestamps order as signed values,
// floating point by the IEEE total order (-NaN < -inf < ... < -0 < +0 < ... < +inf < +NaN; on NaN-free data that is
// the `<` order the reference's comparator uses, sqls_rtti_comp.hpp:84-175).  Masks are ignored (the callers reject
// them like the reference does).  n < 2^32.
gdf_error sort_permutation(const gdf_column* const* cols, int ncols, size_t n, uint32_t* d_perm);

// data[i] = data[perm[i]] for i in [0, n), elements of `width` bytes (1, 2, 4, 8), through a scratch copy.
gdf_error permute_in_place(void* data, int width, size_t n, const uint32_t* d_perm);

// out[i] = in[perm[i]]
gdf_error gather_rows(const void* in, void* out, int width, size_t n, const uint32_t* d_perm);

}  // namespace b200
