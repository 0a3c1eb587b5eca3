#!/usr/bin/env bash
# filter ticket A/B, ncu of the stencil select, full-config reference arm, one-pass exchange emulation test
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_dist_gpu.py tests/test_filter_gpu.py -m gpu -x -q > $OUT/t_dist.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/t_dist.log
timeout 300 python bench.py --only filter --no-e2e --no-cpu > $OUT/filt_ship.json 2> $OUT/filt_ship.err; echo "ship rc=$?"
B200_SELECT_STATIC=1 timeout 300 python bench.py --lab --only filter --no-e2e --no-cpu > $OUT/filt_static.json 2> $OUT/filt_static.err; echo "static rc=$?"
B200_SELECT_STATIC=0 timeout 300 python bench.py --lab --only filter --no-e2e --no-cpu > $OUT/filt_ticket.json 2> $OUT/filt_ticket.err; echo "ticket rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'select_chunked|compare' -c 2 -f -o $OUT/prof_stencil \
   python bench.py --only filter_stencil --steps 1 --warmup 0 --no-e2e --no-cpu > $OUT/prof_stencil.log 2>&1; echo "ncu rc=$?"
timeout 1500 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref rc=$?"; tail -5 $OUT/bench_ref.err
ls -la $OUT | head -30
