#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY - not part of the product path.
#
# Builds the *unmodified algorithmic* reference (gpuopenanalytics/libgdf @ de7b160c) for sm_100a
# into oracle/_ref/ so GPU parity tests and `bench.py --impl reference` can run the reference's own
# kernels next to ours.  The reference tree is read where it lies (/root/reference); the only files
# written are under oracle/_ref/ (git-ignored, but shipped to the GPU box by gpurun).
#
# The reference targets CUDA 9.2 / Thrust 1.9; eight mechanical source fixes are needed for
# CUDA 12.9 / CCCL 2.8.  They are applied with sed to a scratch copy under
# oracle/_ref/src (never committed): none changes an algorithm.
#   1. cudaPointerAttributes.isManaged   -> (type == cudaMemoryTypeManaged)      (field removed in CUDA 11)
#   2. cub::ShuffleIndex(x,0,32,mask)    -> cub::ShuffleIndex<32>(x,0,mask)      (signature change)
#   3. missing <thrust/iterator/constant_iterator.h>, <thrust/host_vector.h> includes
#   4. `const Iterator begin;` member    -> non-const (deleted copy-assign under new Thrust)
#   5. moderngpu __shfl_up/down/__ballot -> *_sync variants (only pulled in by the sort-join header)
# Symbols are NOT renamed; load the result in a process that has not loaded our own libgdf.so,
# or dlopen it with RTLD_LOCAL (tests/_ref_loader does the latter through ctypes).
set -euo pipefail
REF=${REF:-/root/reference/libgdf}
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
if [ ! -d "$REF" ]; then
  echo "build_ref: $REF not present (GPU box?) - keeping prebuilt $OUT" >&2
  exit 0
fi
SRC="$OUT/src"
rm -rf "$SRC"; mkdir -p "$SRC" "$OUT/obj"
cp -r "$REF/src" "$SRC/src"
cp -r "$REF/include" "$SRC/include"
mkdir -p "$SRC/thirdparty"
cp -r "$REF/thirdparty/moderngpu" "$SRC/thirdparty/moderngpu"
cp -r "$REF/thirdparty/cnmem" "$SRC/thirdparty/cnmem"
rm -rf "$SRC/src/tests"

# --- mechanical toolchain patches (see header) ---
sed -i 's/hashtbl_values_ptr_attributes\.isManaged/(hashtbl_values_ptr_attributes.type == cudaMemoryTypeManaged)/' \
  "$SRC/src/hashmap/concurrent_unordered_map.cuh" "$SRC/src/hashmap/concurrent_unordered_multimap.cuh"
sed -i 's/cub::ShuffleIndex(output_offset, 0, warp_size, activemask)/cub::ShuffleIndex<warp_size>(output_offset, 0, activemask)/' \
  "$SRC/src/join/hash/join_kernels.cuh"
for f in filterops.cu streamcompactionops.cu; do
  sed -i '0,/#include <thrust\/functional.h>/s//#include <thrust\/functional.h>\n#include <thrust\/iterator\/constant_iterator.h>\n#include <thrust\/host_vector.h>\n#include <thrust\/device_vector.h>/' "$SRC/src/$f"
done
sed -i 's/^\tconst Iterator begin;/\tIterator begin;/' "$SRC/src/streamcompactionops.cu" "$SRC/src/filterops.cu"
MG="$SRC/thirdparty/moderngpu/src/moderngpu"
sed -i 's/__shfl_up(u\.x\[i\], offset, width)/__shfl_up_sync(0xffffffff, u.x[i], offset, width)/; s/__shfl_down(u\.x\[i\], offset, width)/__shfl_down_sync(0xffffffff, u.x[i], offset, width)/' "$MG/intrinsics.hxx"
sed -i 's/"shfl\."#dir"\.b32 r0|p, %1, %2, %3;"/"shfl.sync."#dir".b32 r0|p, %1, %2, %3, 0xffffffff;"/; s/"shfl\."#dir"\.b32 lo|p, lo, %2, %3;"/"shfl.sync."#dir".b32 lo|p, lo, %2, %3, 0xffffffff;"/; s/"shfl\."#dir"\.b32 hi  , hi, %2, %3;"/"shfl.sync."#dir".b32 hi  , hi, %2, %3, 0xffffffff;"/' "$MG/intrinsics.hxx"
sed -i 's/__ballot(x)/__ballot_sync(0xffffffff, x)/' "$MG/cta_scan.hxx"
sed -i 's/__ballot(has_head_flag)/__ballot_sync(0xffffffff, has_head_flag)/; s/__ballot(0 != storage\.delta\[tid\])/__ballot_sync(0xffffffff, 0 != storage.delta[tid])/' "$MG/cta_segscan.hxx"

NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS=(-std=c++17 --expt-extended-lambda --expt-relaxed-constexpr -O3
       -gencode arch=compute_100a,code=sm_100a -DHASH_JOIN -DHT_LEGACY_ALLOCATOR -w
       -I "$SRC/include" -I "$SRC/src" -I "$MG/.." -I "$SRC/thirdparty/cnmem/include"
       -I /usr/local/cuda/include/nvtx3 -Xcompiler -fPIC)

# hot-path translation units only; everything else is out of scope.
CU=(join/joining sqls_ops hashing streamcompactionops filterops bitmaskops reductions binaryops
    validops cudautils)
CPP=(column context errorhandling nvtx_utils memory/memory memory/memory_manager)

pids=()
compile() { # src obj
  if [ ! -f "$2" ] || [ "$1" -nt "$2" ]; then "$NVCC" "${FLAGS[@]}" -c "$1" -o "$2"; fi
}
JOBS=${JOBS:-6}
running=0
for f in "${CU[@]}"; do
  compile "$SRC/src/$f.cu" "$OUT/obj/$(echo "$f" | tr / _).o" &
  running=$((running+1)); if [ "$running" -ge "$JOBS" ]; then wait -n; running=$((running-1)); fi
done
for f in "${CPP[@]}"; do
  compile "$SRC/src/$f.cpp" "$OUT/obj/$(echo "$f" | tr / _).o" &
  running=$((running+1)); if [ "$running" -ge "$JOBS" ]; then wait -n; running=$((running-1)); fi
done
compile "$SRC/thirdparty/cnmem/src/cnmem.cpp" "$OUT/obj/cnmem.o" &
wait

# one self-contained library: gdf_* and rmm* together (avoids the librmm.so soname clash with ours)
"$NVCC" -shared -o "$OUT/libgdf_ref.so" "$OUT"/obj/*.o -lcuda -Xlinker -Bsymbolic
rm -rf "$OUT/obj" "$SRC"
echo "build_ref: wrote $OUT/libgdf_ref.so"
