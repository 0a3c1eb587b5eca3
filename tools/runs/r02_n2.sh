#!/usr/bin/env bash
# multi-GPU bench as the driver launches it; usage: r02_nN.sh N [extra bench args]
set -u
N=${1:-2}; shift || true
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo_n$N.txt 2>&1
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
   bench.py --gpus $N --steps 5 --warmup 3 "$@" > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err
echo "bench N=$N rc=$?"; tail -c 600 $OUT/bench_n$N.json; tail -5 $OUT/bench_n$N.err
