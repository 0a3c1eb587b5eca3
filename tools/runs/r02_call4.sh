#!/usr/bin/env bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_filter_gpu.py tests/test_reference_parity.py tests/test_arrow.py -m gpu -x -q > $OUT/t_filt.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/t_filt.log
timeout 300 python bench.py --only filter,filter_stencil --no-e2e --no-cpu > $OUT/filt_ship.json 2> $OUT/filt_ship.err; echo "ship rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'select_chunked' -c 1 -f -o $OUT/prof_stencil2 \
   python bench.py --only filter_stencil --steps 1 --warmup 0 --no-e2e --no-cpu > $OUT/prof_stencil2.log 2>&1; echo "ncu rc=$?"
