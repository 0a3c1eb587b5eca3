#!/usr/bin/env bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_join_gpu.py tests/test_reference_parity.py -m gpu -x -q > $OUT/t_join.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/t_join.log
timeout 200 python bench.py --only join --no-e2e --no-cpu > $OUT/join_ovl.json 2> $OUT/join_ovl.err; echo "ship rc=$?"; python tools/show_bench.py $OUT/join_ovl.json
for v in 3 1; do
B200_OVERLAP_SCATTER_CTAS=$v timeout 200 python bench.py --lab --only join --no-e2e --no-cpu > $OUT/join_ovl_c$v.json 2> $OUT/join_ovl_c$v.err; echo "lab ctas=$v"; python tools/show_bench.py $OUT/join_ovl_c$v.json | tail -2
done
B200_OVERLAP_SCATTER_CTAS=2 timeout 200 python bench.py --lab --only join --no-e2e --no-cpu > $OUT/join_ovl_c2.json 2>/dev/null; echo "lab ctas=2"; python tools/show_bench.py $OUT/join_ovl_c2.json | tail -2
B200_BUILD_OVERLAP=0 timeout 200 python bench.py --lab --only join --no-e2e --no-cpu > $OUT/join_ovl_off.json 2>/dev/null; echo "lab overlap off"; python tools/show_bench.py $OUT/join_ovl_off.json | tail -2
