#!/usr/bin/env bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_join_gpu.py tests/test_dist_gpu.py -m gpu -x -q > $OUT/t_join.log 2>&1; echo "pytest rc=$?"; tail -2 $OUT/t_join.log
timeout 200 python bench.py --only join --no-e2e --no-cpu > $OUT/join_db.json 2> $OUT/join_db.err; echo "ship rc=$?"; python tools/show_bench.py $OUT/join_db.json
