#!/usr/bin/env bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_filter_gpu.py tests/test_reference_parity.py tests/test_join_gpu.py -m gpu -x -q -s -k "not matrix" > $OUT/t_filt.log 2>&1; echo "pytest rc=$?"; grep "gdf_filter returned" $OUT/t_filt.log; tail -3 $OUT/t_filt.log
timeout 300 python bench.py --only filter,filter_stencil --no-e2e --no-cpu > $OUT/filt_ship.json 2> $OUT/filt_ship.err; echo "ship rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'select_chunked' -c 1 -f -o $OUT/prof_stencil3 \
   python bench.py --only filter_stencil --steps 1 --warmup 0 --no-e2e --no-cpu > $OUT/prof_stencil3.log 2>&1; echo "ncu rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'probe32' -c 1 -f -o $OUT/prof_probe \
   python bench.py --only join --steps 1 --warmup 0 --no-e2e --no-cpu > $OUT/prof_probe.log 2>&1; echo "ncu probe rc=$?"
