#!/usr/bin/env bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_dist_gpu.py tests/test_join_gpu.py -m gpu -x -q > $OUT/t_join.log 2>&1; echo "pytest rc=$?"; tail -12 $OUT/t_join.log | cut -c1-200
