#!/usr/bin/env bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 400 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $OUT/pytest_gpu.log
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/bench_b200.json 2> $OUT/bench_b200.err; echo "bench rc=$?"
for w in join groupby filter filter_stencil reduce_sum add_i64_2.5e8 hash_partition join_result_cols; do
  timeout 120 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
     -k regex:'probe32|part_|build32|fixup|build_fast|extract_fast|select_|compare_static|reduce_kernel|binary_|partition_|gather_kernel' \
     --csv --log-file $OUT/traffic_$w.csv python bench.py --only $w --steps 1 --warmup 0 --no-e2e --no-cpu > $OUT/traffic_$w.log 2>&1
  echo "traffic $w rc=$?"
done
