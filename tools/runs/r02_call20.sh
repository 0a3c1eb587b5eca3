#!/usr/bin/env bash
set -u
OUT=gpurun_out; mkdir -p $OUT
B200_SCATTER_BULK=1 timeout 200 python bench.py --lab --only join --no-e2e --no-cpu > $OUT/join_bulk.json 2> $OUT/join_bulk.err; echo "bulk rc=$?"; python tools/show_bench.py $OUT/join_bulk.json
B200_SCATTER_BULK=0 timeout 200 python bench.py --lab --only join --no-e2e --no-cpu > $OUT/join_nobulk.json 2> $OUT/join_nobulk.err; echo "nobulk rc=$?"; python tools/show_bench.py $OUT/join_nobulk.json
timeout 300 python bench.py --only filter_stencil,groupby --no-e2e --no-cpu > $OUT/misc.json 2> $OUT/misc.err; python tools/show_bench.py $OUT/misc.json
