#!/usr/bin/env bash
set -u
OUT=gpurun_out; mkdir -p $OUT
for v in 20 19; do
B200_ROWS_PER_PART_LOG2=$v timeout 200 python bench.py --lab --only join --no-e2e --no-cpu > $OUT/join_rpp$v.json 2> $OUT/join_rpp$v.err; echo "rpp $v rc=$?"; python tools/show_bench.py $OUT/join_rpp$v.json
done
