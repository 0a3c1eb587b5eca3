#!/usr/bin/env bash
# final profiling pass (1 GPU): DRAM traffic of every kernel at full size + ncu --set full of the kernels without a capture yet
set -u
OUT=gpurun_out; mkdir -p $OUT
bash tools/traffic_full.sh 2>&1 | tail -10
cap() { name=$1; regex=$2; wl=$3
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$regex -c 1 -f -o $OUT/prof_$name \
     python bench.py --only $wl --steps 1 --warmup 0 --no-e2e --no-cpu > $OUT/prof_$name.log 2>&1; echo "ncu $name rc=$?"; }
cap scatter 'part_scatter32_bulk' join
cap build 'build32_kernel' join
cap reduce 'reduce_kernel' reduce_sum
cap binary 'binary_' add_i64_2.5e8
cap hashpart 'partition_scatter_kernel' hash_partition
cap gather 'gather_kernel' join_result_cols
ls -la $OUT | grep -c prof_
