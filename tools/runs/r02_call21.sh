#!/usr/bin/env bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_gpu.log
timeout 300 python bench.py --only join,c5 --no-e2e --no-cpu > $OUT/join_c5.json 2> $OUT/join_c5.err; echo "rc=$?"; python tools/show_bench.py $OUT/join_c5.json
