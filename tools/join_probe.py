"""Diagnostic: where does a join step spend its time? (host wall clock around call / free)"""
import sys, time, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.1
api = bench.Api("b200")
wl = bench.JoinWorkload(api, scale)
api.profile_begin()

def loop(tag, n=4, sync=True):
    for i in range(n):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        wl.call(wl.probe, wl.build)
        if sync: torch.cuda.synchronize()
        t1 = time.perf_counter()
        wl.free()
        if sync: torch.cuda.synchronize()
        t2 = time.perf_counter()
    torch.cuda.synchronize()
    prof = api.profile_end()
    print("%s: last call %.2f ms free %.2f ms kernels/step %.2f ms" % (tag, (t1 - t0) * 1e3, (t2 - t1) * 1e3, sum(v["ms"] for v in prof.values()) / n))

loop("A sync each step")
loop("B no sync", sync=False)
ms, _ = bench.timed_steps(wl.step, 1, 4); api.profile_end()
print("B2 timed_steps: %.2f ms/step" % ms)
clocks = bench.Clocks(0)
time.sleep(0.3)
ms, _ = bench.timed_steps(wl.step, 1, 4); api.profile_end()
print("C with NVML poll thread: %.2f ms/step" % ms)
clocks.stop(); time.sleep(0.2)
ms, _ = bench.timed_steps(wl.step, 1, 4); api.profile_end()
print("C2 poll thread stopped: %.2f ms/step" % ms)
ok = wl.check()
ms, _ = bench.timed_steps(wl.step, 1, 4); api.profile_end()
print("D after check(): %.2f ms/step (check ok=%s)" % (ms, ok))
torch.cuda.empty_cache()
ms, _ = bench.timed_steps(wl.step, 1, 4); api.profile_end()
print("E after empty_cache: %.2f ms/step" % ms)
