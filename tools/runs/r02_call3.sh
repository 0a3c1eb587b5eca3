#!/usr/bin/env bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_dist_gpu.py tests/test_filter_gpu.py tests/test_reference_parity.py -m gpu -x -q > $OUT/t_filt.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/t_filt.log
timeout 300 python bench.py --only filter,filter_stencil --no-e2e --no-cpu > $OUT/filt_ship.json 2> $OUT/filt_ship.err; echo "ship rc=$?"
B200_SELECT_STATIC=1 timeout 300 python bench.py --lab --only filter --no-e2e --no-cpu > $OUT/filt_static.json 2> $OUT/filt_static.err; echo "static rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'build_fast' -c 1 -f -o $OUT/prof_groupby \
   python bench.py --only groupby --steps 1 --warmup 0 --no-e2e --no-cpu > $OUT/prof_groupby.log 2>&1; echo "ncu gb rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'select_stream' -c 1 -f -o $OUT/prof_filter \
   python bench.py --only filter --steps 1 --warmup 0 --no-e2e --no-cpu > $OUT/prof_filter.log 2>&1; echo "ncu filter rc=$?"
ls -la $OUT | head -30
