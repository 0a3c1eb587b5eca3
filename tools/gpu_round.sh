#!/usr/bin/env bash
# One gpurun call: GPU parity tests, both bench arms, ncu launch lists and full captures of the top kernels.
# Usage (from the repo root, under gpurun):  bash tools/gpu_round.sh [stages]   stages default "test bench ref list full"
set -u
STAGES=${1:-"test bench ref list full"}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi > $OUT/smi.txt 2>&1
has() { [[ " $STAGES " == *" $1 "* ]]; }

if has test; then
  timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1
  echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
fi
if has smoke; then
  timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1
  echo "smoke rc=$?"; tail -2 $OUT/smoke.log
fi
if has bench; then
  timeout 900 python bench.py > $OUT/bench_b200.json 2> $OUT/bench_b200.err
  echo "bench rc=$?"; cat $OUT/bench_b200.json; tail -5 $OUT/bench_b200.err
fi
if has ref; then
  timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err
  echo "ref rc=$?"; cat $OUT/bench_ref.json; tail -5 $OUT/bench_ref.err
fi
if has list; then
  for w in join groupby filter; do
    timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_$w.csv \
      python bench.py --only $w --steps 2 --warmup 1 --no-e2e --no-cpu > $OUT/launches_$w.log 2>&1
    echo "list $w rc=$?"
  done
fi
if has traffic; then
  # DRAM bytes of every library kernel at FULL size (single-pass counters, no kernel replay) -> gpurun_out/r02_traffic_full_size.json
  bash tools/traffic_full.sh
fi
if has full; then
  # quarter-size inputs: ncu's kernel replay saves/restores every written allocation; the partition
  # geometry (rows per partition, table bytes per partition) is the same as at full size.
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'part_|select_|build_fast|extract_fast' \
    -c 24 -f -o $OUT/prof_full python bench.py --scale 0.25 --steps 1 --warmup 0 --no-e2e --no-cpu > $OUT/prof_full.log 2>&1
  echo "full rc=$?"
fi
ls -la $OUT
