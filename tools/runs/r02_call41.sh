#!/usr/bin/env bash
set -u
N=${1:-2}
OUT=gpurun_out; mkdir -p $OUT
run() { tag=$1; shift
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
     bench.py --gpus $N --steps 5 --warmup 3 --only join --no-e2e "$@" > $OUT/sweep_n${N}_$tag.json 2> $OUT/sweep_n${N}_$tag.err
  echo "$tag rc=$?"; python tools/show_bench.py $OUT/sweep_n${N}_$tag.json | head -3; grep -i "error\|Traceback" $OUT/sweep_n${N}_$tag.err | head -5
  python - <<PY
import json
try:
    d=json.loads(open('$OUT/sweep_n${N}_$tag.json').read().strip().splitlines()[-1])
    print({k: round(v['ms_per_step'],3) for k,v in d['kernels_rank0'].items()}, d['parity_properties_ok'])
except Exception as e: print('no result', e)
PY
}
run even
run even3 --scatter-ctas 0
