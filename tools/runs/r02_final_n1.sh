#!/usr/bin/env bash
# the driver's end-of-round sequence at N=1: GPU tests, smoke, reference arm, bench - with wall times
set -u
OUT=gpurun_out; mkdir -p $OUT
t0=$(date +%s)
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$? $(( $(date +%s)-t0 ))s"; tail -2 $OUT/pytest_gpu.log
t0=$(date +%s); timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$? $(( $(date +%s)-t0 ))s"; tail -1 $OUT/smoke.log
t0=$(date +%s); timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref rc=$? $(( $(date +%s)-t0 ))s"; python tools/show_bench.py $OUT/bench_ref.json
t0=$(date +%s); timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/bench_b200.json 2> $OUT/bench_b200.err; echo "bench rc=$? $(( $(date +%s)-t0 ))s"; python tools/show_bench.py $OUT/bench_b200.json; tail -3 $OUT/bench_b200.err
