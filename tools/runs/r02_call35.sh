#!/usr/bin/env bash
set -u
OUT=gpurun_out; mkdir -p $OUT
for v in 1048576 96 64 32; do
B200_BUILD_ROUND_MB=$v timeout 200 python bench.py --lab --only join --no-e2e --no-cpu > $OUT/join_round$v.json 2> $OUT/join_round$v.err; echo "lab round MB=$v"; python tools/show_bench.py $OUT/join_round$v.json | tail -2
done
