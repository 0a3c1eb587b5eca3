"""Lab (2+ GPUs, torchrun): time the fused-exchange scatter alone - all ranks at once vs one rank at a time - to see
whether a busy receiver is what holds the exchange at 0.43 TB/s per GPU (profiles/r02_notes.md 5c)."""
import os, sys, time
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from libgdf_b200 import dist as D

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
ops = D.GdfOps()
peer = D.PeerExchange(ops)
P = int(1e9) // world
g = torch.Generator(device="cuda"); g.manual_seed(1000 * (rank + 1))
probe = torch.randint(0, int(1e8), (P,), generator=g, device="cuda", dtype=torch.int64)
nlocal = D.fused_nlocal(int(1e8), world)
bins = world * nlocal
cp, _ = ops.xjoin_count(probe, world, nlocal)
mine = torch.tensor(cp, dtype=torch.int64, device="cuda")
allc = torch.empty(world * bins, dtype=torch.int64, device="cuda")
dist.all_gather_into_tensor(allc, mine)
M = allc.view(world, bins).cpu().tolist()
off, tot, recv = D.plan_fused_exchange(M, world, nlocal, rank)
slot = peer._ensure_pairs("xprobe", max(recv))
flag = torch.zeros(1, dtype=torch.int32, device="cuda")

def timed(active):
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        if active:
            ops.xjoin_scatter(probe, 0, world, nlocal, slot["peers"][0], off)
    e1.record()
    torch.cuda.synchronize(); dist.barrier()
    return e0.elapsed_time(e1) / 3

for mode in ["all"] + ["only%d" % r for r in range(min(world, 2))]:
    active = mode == "all" or mode == "only%d" % rank
    timed(active)
    ms = timed(active)
    remote_gb = 8 * P * (world - 1) / world / 1e9
    print("rank %d mode %-6s scatter %.3f ms -> %.0f GB/s of remote pairs" % (rank, mode, ms, remote_gb / (ms * 1e-3) if active else 0), flush=True)
peer.close()
dist.destroy_process_group()
