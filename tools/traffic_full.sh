#!/usr/bin/env bash
# Under gpurun: DRAM traffic of every library kernel at full size, one bench step per workload (single-pass counters, no replay).
# Result: gpurun_out/traffic_<workload>.csv ; then, in the container:  see the last line.
set -u
OUT=gpurun_out; mkdir -p $OUT
for w in join groupby filter filter_stencil reduce_sum add_i64_2.5e8 hash_partition join_result_cols; do
  timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
     -k regex:'probe32|part_|build32|fixup|build_fast|extract_fast|select_|compare_static|reduce_kernel|binary_|partition_|gather_kernel' \
     --csv --log-file $OUT/traffic_$w.csv python bench.py --only $w --steps 1 --warmup 0 --no-e2e --no-cpu > $OUT/traffic_$w.log 2>&1
  echo "traffic $w rc=$?"
done
python tools/traffic_from_ncu.py $(for w in join groupby filter filter_stencil reduce_sum add_i64_2.5e8 hash_partition join_result_cols; do echo $w=$OUT/traffic_$w.csv; done) > $OUT/r02_traffic_full_size.json && echo "wrote $OUT/r02_traffic_full_size.json (copy to profiles/)"
