#!/usr/bin/env python
"""BASELINE config C5: gdf_left_join on a composite (int64,int32) key with 30 % null rows, 5e8 x 5e7 rows,
N GPUs (torchrun) - or the single-GPU gdf_left_join call when N = 1.  Prints one JSON line on rank 0.

  python tools/bench_c5.py --steps 3                                   (1 GPU)
  python -m torch.distributed.run --nproc-per-node 8 ... tools/bench_c5.py --steps 3

Synthetic input as in SURVEY.md 8(d): key0 int64 uniform [0,5e7), key1 int32 uniform [0,4); 30 % of the rows of
each side are null rows, the null bit cleared in key0's or key1's mask with equal probability.
Parity at full size through properties (all-reduced over the ranks): every left row appears at least once,
null left rows appear exactly once and with -1, every matched pair joins equal keys of two valid rows."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def make_side(n, key_range, seed, dev):
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    k0 = torch.randint(0, key_range, (n,), generator=g, device=dev, dtype=torch.int64)
    k1 = torch.randint(0, 4, (n,), generator=g, device=dev, dtype=torch.int32)
    null_row = torch.rand(n, generator=g, device=dev) < 0.3
    which = torch.rand(n, generator=g, device=dev) < 0.5
    v0, v1 = ~(null_row & which), ~(null_row & ~which)
    return k0, k1, v0, v1


def pack(bits):  # bool[n] -> Arrow LSB-first bitmask (test-data preparation, not the product path)
    n = bits.numel()
    pad = (-n) % 8
    b = torch.cat([bits, torch.zeros(pad, dtype=torch.bool, device=bits.device)]).view(-1, 8).to(torch.uint8)
    w = torch.tensor([1, 2, 4, 8, 16, 32, 64, 128], dtype=torch.uint8, device=bits.device)
    return (b * w).sum(1).to(torch.uint8)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--scale", type=float, default=1.0)
    args = ap.parse_args()
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", ""):
        os.environ["NCCL_DEBUG"] = "WARN"
    dist.init_process_group("nccl" if world > 1 else "gloo", device_id=dev if world > 1 else None,
                            init_method=None if world > 1 else "tcp://127.0.0.1:29577", rank=rank, world_size=world)
    from libgdf_b200.librmm_cffi import librmm, librmm_config
    librmm_config.use_pool_allocator = True
    librmm.finalize()
    librmm.initialize()
    from libgdf_b200 import dist as D
    ops = D.GdfOps()
    NL, NR = int(5e8 * args.scale), int(5e7 * args.scale)
    plo, phi = D.shard_bounds(NL, world, rank)
    blo, bhi = D.shard_bounds(NR, world, rank)
    l0, l1, lv0, lv1 = make_side(phi - plo, NR, 100 + rank, dev)
    r0, r1, rv0, rv1 = make_side(bhi - blo, NR, 200 + rank, dev)
    lmask, rmask = [pack(lv0), pack(lv1)], [pack(rv0), pack(rv1)]

    def step():
        return D.distributed_left_join_masked([l0, l1], lmask, [r0, r1], rmask, plo, blo, ops)

    # ---- parity properties ----
    gl, gr = step()
    glong, matched = gl.long(), gr >= 0
    seen = torch.zeros(NL, dtype=torch.int32, device=dev)
    seen.index_add_(0, glong, torch.ones_like(gl))
    if world > 1:
        dist.all_reduce(seen)
    lvalid_local = lv0 & lv1
    mine = seen[plo:phi]
    ok = bool((mine >= 1).all()) and bool((mine[~lvalid_local] == 1).all())

    def gathered(x, total, lo, hi):
        if world == 1:
            return x
        full = torch.empty(total, dtype=x.dtype, device=dev)
        dist.all_gather([full[a:b] for a, b in (D.shard_bounds(total, world, k) for k in range(world))], x)
        return full
    L0, L1, LV = gathered(l0, NL, plo, phi), gathered(l1, NL, plo, phi), gathered(lvalid_local, NL, plo, phi)
    R0, R1, RV = gathered(r0, NR, blo, bhi), gathered(r1, NR, blo, bhi), gathered(rv0 & rv1, NR, blo, bhi)
    ml, mr = glong[matched], gr[matched].long()
    ok = ok and bool((L0[ml] == R0[mr]).all()) and bool((L1[ml] == R1[mr]).all()) and bool(LV[ml].all()) and bool(RV[mr].all())
    pairs = torch.tensor([gl.numel()], dtype=torch.int64, device=dev)
    flag = torch.tensor([int(ok)], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(pairs)
        dist.all_reduce(flag)
    del seen, L0, L1, LV, R0, R1, RV, gl, gr, glong
    torch.cuda.empty_cache()

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    timings = {}
    e0.record()
    for _ in range(args.steps):
        a, b = D.distributed_left_join_masked([l0, l1], lmask, [r0, r1], rmask, plo, blo, ops, timings=timings)
        del a, b
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / args.steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"metric": "rows_per_sec_left_join_composite_key_nulls", "value": (NL + NR) / (ms.item() * 1e-3),
                          "unit": "rows/s", "n_gpus": world, "ms_per_step": ms.item(), "steps": args.steps,
                          "config": {"workload": "C5 gdf_left_join (int64,int32) key, 30 %% null rows, %d x %d" % (NL, NR)},
                          "output_pairs": int(pairs.item()), "parity_properties_ok": int(flag.item()) == world,
                          "phases_ms_rank0": {k: v / args.steps for k, v in D.resolve_timings(timings).items()}}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
