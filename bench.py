#!/usr/bin/env python
"""Benchmark of the gdf hot path on B200.

Exactly ONE JSON line on stdout (anything a library prints to fd 1 is diverted to stderr).

N = 1   headline = BASELINE.json's metric "rows/sec hash inner-join ... int64": C3 = gdf_inner_join of 1e9
        probe rows x 1e8 build rows, int64 keys (build = permutation, probe uniform, every probe row matches
        once).  A step = one gdf_inner_join call over device-resident columns + gdf_column_free of its outputs.
        The same line carries, under `workloads` and each with its own roofline object and full-size parity
        properties: C4 gdf_group_by_sum (1e9 rows, 1e6 int64 groups, Zipf s=1.05; uniform-key variants), C2
        gdf_filter (1e9 int64 rows, 10 % selectivity) and its gpu_comparison_static + gpu_apply_stencil variant,
        gdf_sum_i64 (plain / masked), gdf_add (C1 and 2.5e8 int64 rows), gdf_hash_partition, a join with
        result_cols, and C5 (left join, composite key, 30 % null rows) on one GPU.
  value      rows/s (probe+build rows per step / CUDA-event time), inputs resident in HBM
  e2e        the same C-ABI call from pinned HOST buffers as ONE call: H2D of both key columns and D2H of both
             index columns inside the timed region (PCIe-bound); e2e.pipelined = the probe column streamed
             through the same call in 8 pieces with the copies overlapped
  roofline   dominant kernel of the headline step: algorithmic bytes of that kernel / its average launch
             duration (CUDA events recorded by the library around every launch, live in the timed steps)
             vs MEASURED_PEAKS.json hbm_gbs; traffic = DRAM bytes per launch from the full-size ncu capture
             (profiles/r02_traffic_full_size.json, reported only while the SHA-1 of csrc/ it was captured from
             matches the sources: tools/traffic_full.sh regenerates it)
  cpu_baseline  the C oracle port (single thread) on a bounded sample of C3; cpu_baseline_pyarrow beside it
N > 1   (torchrun) strong scaling of C3: the two tables are block-distributed over the ranks; every step
        partitions both sides by (destination rank x the receiver's local partition) INSIDE one kernel that
        stores the rows into the peers' partition-contiguous receive buffers over NVLink (--exchange p2p,
        default; hash-partition + NCCL all_to_all when CUDA IPC peer mapping is unavailable or with
        --exchange nccl), then builds and probes locally.  Time = max over ranks of the CUDA-event time between
        two barriers.  The same launch then times C4 (two-phase distributed group-by) and C5 (distributed left
        join on the composite key with NULLs) under `workloads`, with per-phase device times.
--impl reference   the reference's own kernels (oracle/_ref/libgdf_ref.so = gpuopenanalytics/libgdf rebuilt for
        sm_100a; the reference is a CUDA library and has no CPU implementation) through the identical harness
        on the SAME config: C3 at 1e9 x 1e8 and C2 at 1e9 rows.  Only its group-by is bounded (5e8 rows with
        uniform keys - its int arithmetic overflows above 2^29 rows - and a 5e6-row Zipf sample: its CAS loop
        serialises on the hot key).  Under torchrun only rank 0 runs it.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

SEED = 0xabcdef  # reference python/tests/utils.py:31-33


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--scale", type=float, default=1.0, help="shrink every row count (debug only)")
    ap.add_argument("--only", default="", help="comma list of workload keys (join, groupby, groupby_uniform, groupby_all, filter, "
                    "filter_stencil, reduce_sum, add_c1, hash_partition, join_result_cols, c5, ...; debug only)")
    ap.add_argument("--exchange", choices=["p2p", "nccl"], default="p2p",
                    help="N>1: rows cross NVLink inside the partition kernel (peer memory) or via NCCL all_to_all")
    ap.add_argument("--lab", action="store_true", help="load the -DB200_LAB build (libgdf_b200/lib_lab, make LAB=1): "
                    "environment-variable knobs and timing ablations; never a bench value")
    ap.add_argument("--xjoin-rpp-log2", type=int, default=0, help="N > 1: log2 of the build rows per receiver-local partition of "
                    "the one-pass exchange (0 = library default)")
    ap.add_argument("--no-overlap", action="store_true", help="N > 1: fill the hash tables after the exchange instead of beside it")
    ap.add_argument("--scatter-ctas", type=int, default=-1, help="N > 1: CTAs per SM of the probe-side scatter while the tables are built (0 = all)")
    ap.add_argument("--symm-mem", action="store_true", help="N > 1: receive buffers from torch symmetric memory (VMM) instead of CUDA IPC")
    ap.add_argument("--two-pass", action="store_true", help="N > 1: round 1's exchange (partition by destination, then again locally)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# library loading: ours through the package, the reference through the same cdef
# ------------------------------------------------------------------------------------------------
class Api(object):
    def __init__(self, impl, lab=False):
        self.impl = impl
        if impl == "b200":
            if lab:
                import libgdf_b200
                libgdf_b200.LIB_DIR = os.path.join(ROOT, "libgdf_b200", "lib_lab")
            from libgdf_b200.librmm_cffi import librmm, librmm_config
            librmm_config.use_pool_allocator = True
            librmm.finalize()
            librmm.initialize()
            from libgdf_b200.libgdf_cffi import ffi, libgdf_api
            self.ffi, self.lib = ffi, libgdf_api
            self.profile = True
        else:
            import cffi
            from libgdf_b200._cdef import header_cdef
            path = os.path.join(ROOT, "oracle", "_ref", "libgdf_ref.so")
            if not os.path.isfile(path):
                raise FileNotFoundError(path)
            ffi = cffi.FFI()
            ffi.cdef(header_cdef("gdf/cffi/types.h", "gdf/cffi/functions.h", "memory.h"))
            self.ffi, self.lib = ffi, ffi.dlopen(path)
            opts = ffi.new("rmmOptions_t*")
            opts.allocation_mode = self.lib.PoolAllocation   # the reference's own test config (conftest.py:7-9)
            opts.initial_pool_size = 48 << 30
            opts.enable_logging = False
            rc = self.lib.rmmInitialize(opts)
            if rc != 0:
                raise RuntimeError("reference rmmInitialize -> %d" % rc)
            self.profile = False

    def check(self, rc, what):
        if rc != 0:
            name = self.ffi.string(self.lib.gdf_error_get_name(rc)).decode()
            raise RuntimeError("%s -> %s" % (what, name))

    def column(self, tensor):
        ffi, lib = self.ffi, self.lib
        col = ffi.new("gdf_column*")
        code = {torch.int64: lib.GDF_INT64, torch.int32: lib.GDF_INT32}[tensor.dtype]
        lib.gdf_column_view(col, ffi.cast("void*", tensor.data_ptr()), ffi.NULL, tensor.numel(), code)
        return col

    def profile_begin(self):
        if self.profile:
            self.lib.gdfx_profile_enable(1)
            self.profile_end()

    def profile_end(self):
        if not self.profile:
            return {}
        buf = self.ffi.new("char[]", 1 << 16)
        self.lib.gdfx_profile_report(buf, 1 << 16)
        return json.loads(self.ffi.string(buf).decode())


def alias(ffi, cdata_ptr, n, np_dtype):
    class _A(object):
        pass
    a = _A()
    a.__cuda_array_interface__ = {"shape": (n,), "typestr": np.dtype(np_dtype).str,
                                  "data": (int(ffi.cast("uintptr_t", cdata_ptr)), False), "version": 2, "strides": None}
    return torch.as_tensor(a, device="cuda")


# ------------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md "clocks DURING the timed region")
# ------------------------------------------------------------------------------------------------
class Clocks(object):
    """SM clock + throttle reasons sampled through NVML from a background thread while the timed
    regions run.  (An `nvidia-smi -lms` child process was measured to slow every driver call of this
    process down - a 42 ms join step became 280 ms - so the same counters are read in-process at a
    low rate instead.)"""
    PERIOD_S = 0.05

    def __init__(self, index):
        self.rows, self.stop_flag, self.handle = [], False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[index]) if visible and visible.split(",")[index].isdigit() else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
        except Exception:
            self.handle = None

    def _poll(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)
                reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                self.rows.append((time.perf_counter(), sm, int(reasons)))
            except Exception:
                pass
            time.sleep(self.PERIOD_S)

    def window(self, t0, t1):
        return [r for r in self.rows if t0 <= r[0] <= t1] or self.rows[-2:]

    def stop(self):
        self.stop_flag = True

    def summarise(self, rows):
        if not rows or self.handle is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        nv = self.nv
        bits = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}
        sm = sorted(r[1] for r in rows)
        reasons = [n for n, b in bits.items() if any(r[2] & b for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.max_sm, "reasons": reasons, "samples": len(rows)}


# ------------------------------------------------------------------------------------------------
# workloads
# ------------------------------------------------------------------------------------------------
def gen(seed_offset):
    g = torch.Generator(device="cuda")
    g.manual_seed(SEED + seed_offset)
    return g


def timed_steps(fn, warmup, steps):
    """W untimed + exactly K timed steps, CUDA events on the launching (default) stream."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    return e0.elapsed_time(e1) / steps, (t0, t1)


class JoinWorkload(object):
    name = "C3 gdf_inner_join 1e9 x 1e8 int64 (build = permutation, probe uniform, 100% hit)"

    def __init__(self, api, scale, rank=0):
        self.api = api
        self.B, self.P = int(1e8 * scale), int(1e9 * scale)
        self.build = torch.randperm(self.B, generator=gen(1 + 1000 * rank), device="cuda", dtype=torch.int64)
        self.probe = torch.randint(0, self.B, (self.P,), generator=gen(0 + 1000 * rank), device="cuda", dtype=torch.int64)
        ffi, lib = api.ffi, api.lib
        self.ctx = ffi.new("gdf_context*")
        lib.gdf_context_view(self.ctx, 0, lib.GDF_HASH, 0, 0, 0)
        self.idx = ffi.new("int[]", [0])
        self.out_l, self.out_r = ffi.new("gdf_column*"), ffi.new("gdf_column*")
        self.rows_per_step = self.P + self.B
        self.pairs = None

    def call(self, probe, build):
        ffi, lib = self.api.ffi, self.api.lib
        lc, rc = self.api.column(probe), self.api.column(build)
        la, ra = ffi.new("gdf_column*[]", [lc]), ffi.new("gdf_column*[]", [rc])
        self.api.check(lib.gdf_inner_join(la, 1, self.idx, ra, 1, self.idx, 1, 0, ffi.NULL, self.out_l, self.out_r,
                                          self.ctx), "gdf_inner_join")
        self.pairs = int(self.out_l.size)

    def free(self):
        lib = self.api.lib
        for o in (self.out_l, self.out_r):
            if o.data != self.api.ffi.NULL:
                lib.gdf_column_free(o)

    def step(self):
        self.call(self.probe, self.build)
        self.free()

    def algorithmic_bytes(self):  # key columns in + int32 index pairs out
        return 8 * (self.P + self.B) + 8 * self.pairs

    def check(self):
        """size-independent properties at full size: pair count, every pair joins equal keys,
        left indices are a permutation of the probe rows (100 % hit, unique build keys)."""
        self.call(self.probe, self.build)
        n = self.pairs
        li = alias(self.api.ffi, self.out_l.data, n, np.int32).long()
        ri = alias(self.api.ffi, self.out_r.data, n, np.int32).long()
        ok = n == self.P and bool((self.probe[li] == self.build[ri]).all())
        ok = ok and int(li.sum().item()) == self.P * (self.P - 1) // 2
        del li, ri
        self.free()
        return ok

    def e2e_setup(self):
        self.h_probe = torch.empty(self.P, dtype=torch.int64, pin_memory=True)
        self.h_build = torch.empty(self.B, dtype=torch.int64, pin_memory=True)
        self.h_probe.copy_(self.probe)
        self.h_build.copy_(self.build)
        self.d_probe, self.d_build = torch.empty_like(self.probe), torch.empty_like(self.build)
        self.h_out_l = torch.empty(self.P, dtype=torch.int32, pin_memory=True)
        self.h_out_r = torch.empty(self.P, dtype=torch.int32, pin_memory=True)
        torch.cuda.synchronize()
        return 8 * (self.P + self.B), 8 * self.P

    def e2e_step(self):
        """One host-to-host join through one C-ABI call: H2D both key columns, gdf_inner_join, D2H both index columns."""
        self.d_probe.copy_(self.h_probe, non_blocking=True)
        self.d_build.copy_(self.h_build, non_blocking=True)
        self.call(self.d_probe, self.d_build)
        n = self.pairs
        self.h_out_l[:n].copy_(alias(self.api.ffi, self.out_l.data, n, np.int32), non_blocking=True)
        self.h_out_r[:n].copy_(alias(self.api.ffi, self.out_r.data, n, np.int32), non_blocking=True)
        torch.cuda.synchronize()
        self.free()
        self.e2e_pairs = n

    def e2e_step_pipelined(self, chunks=8):
        """The same host-to-host join the way a caller streams a big probe table through the C ABI: the probe
        column goes up in `chunks` pieces on a copy stream, gdf_inner_join runs per piece against the resident
        build column (default stream), and each piece's index pairs (left ids rebased to the whole table) go
        down on a second copy stream - so H2D, the join and D2H overlap and PCIe is used in both directions at
        once.  Same bytes, same result set as e2e_step (pair order is unspecified in both)."""
        ffi, lib = self.api.ffi, self.api.lib
        P = self.P
        per = ((P + chunks - 1) // chunks + 1) & ~1          # even: 16-byte aligned int64 slices
        if per < self.B:                                       # keep the build side on the right (no INNER flip)
            return self.e2e_step()
        if not hasattr(self, "s_in"):
            self.s_in, self.s_out = torch.cuda.Stream(), torch.cuda.Stream()
        cur = torch.cuda.current_stream()
        bounds = [(lo, min(P, lo + per)) for lo in range(0, P, per)]
        with torch.cuda.stream(self.s_in):
            self.d_build.copy_(self.h_build, non_blocking=True)
            ready = []
            for lo, hi in bounds:
                self.d_probe[lo:hi].copy_(self.h_probe[lo:hi], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.s_in)
                ready.append(ev)
        total, inflight = 0, []
        for (lo, hi), ev in zip(bounds, ready):
            cur.wait_event(ev)
            out_l, out_r = ffi.new("gdf_column*"), ffi.new("gdf_column*")
            lc, rc = self.api.column(self.d_probe[lo:hi]), self.api.column(self.d_build)
            la, ra = ffi.new("gdf_column*[]", [lc]), ffi.new("gdf_column*[]", [rc])
            self.api.check(lib.gdf_inner_join(la, 1, self.idx, ra, 1, self.idx, 1, 0, ffi.NULL, out_l, out_r, self.ctx),
                           "gdf_inner_join")
            n = int(out_l.size)
            if n:
                l, r = alias(ffi, out_l.data, n, np.int32), alias(ffi, out_r.data, n, np.int32)
                if lo:
                    l.add_(lo)                                 # piece-relative -> table-relative left row ids
                done = torch.cuda.Event()
                done.record(cur)
                self.s_out.wait_event(done)
                with torch.cuda.stream(self.s_out):
                    self.h_out_l[total:total + n].copy_(l, non_blocking=True)
                    self.h_out_r[total:total + n].copy_(r, non_blocking=True)
                    copied = torch.cuda.Event()
                    copied.record(self.s_out)
                inflight.append((copied, out_l, out_r))
                total += n
            while len(inflight) > 2:                           # free a piece's outputs once its D2H has finished
                copied, a, b = inflight.pop(0)
                copied.synchronize()
                lib.gdf_column_free(a)
                lib.gdf_column_free(b)
        torch.cuda.synchronize()
        for _, a, b in inflight:
            lib.gdf_column_free(a)
            lib.gdf_column_free(b)
        self.e2e_pairs = total

    def e2e_verify(self):
        """Host-side check of the last e2e result: pair count and the left ids form a permutation of the probe rows."""
        n = self.e2e_pairs
        return n == self.P and int(self.h_out_l[:n].numpy().sum(dtype=np.int64)) == self.P * (self.P - 1) // 2


class GroupbyWorkload(object):
    name = "C4 gdf_group_by_sum 1e9 rows, 1e6 int64 groups, Zipf s=1.05, int64 values"

    def __init__(self, api, scale, rows=None, zipf=True, groups=None):
        self.api = api
        self.N = int((rows or 1e9) * scale)
        self.G = max(int((groups or 1e6) * scale), 16)
        if not zipf:
            self.name = "C4 variant: gdf_group_by_sum %.0e rows, %.0e int64 groups, UNIFORM keys, int64 values" % (self.N, self.G)
        elif rows:
            self.name = "C4 sample: gdf_group_by_sum %.0e rows, %.0e int64 groups, Zipf s=1.05, int64 values" % (self.N, self.G)
        ids = torch.randperm(self.G, generator=gen(3), device="cuda", dtype=torch.int64) * 7919 + 13
        self.keys = torch.empty(self.N, dtype=torch.int64, device="cuda")
        chunk = 1 << 27
        g = gen(2)
        if zipf:
            ranks = torch.arange(1, self.G + 1, device="cuda", dtype=torch.float64)
            cdf = torch.cumsum(ranks.pow(-1.05), 0)
            cdf /= cdf[-1].clone()
        for lo in range(0, self.N, chunk):  # inverse-CDF sampling in chunks (float64 temporaries)
            hi = min(self.N, lo + chunk)
            if zipf:
                u = torch.rand(hi - lo, generator=g, device="cuda", dtype=torch.float64)
                self.keys[lo:hi] = ids[torch.searchsorted(cdf, u).clamp_(max=self.G - 1)]
            else:
                u = torch.randint(0, self.G, (hi - lo,), generator=g, device="cuda", dtype=torch.int64)
                self.keys[lo:hi] = ids[u]
            del u
        self.vals = torch.randint(0, 1000, (self.N,), generator=gen(4), device="cuda", dtype=torch.int64)
        # the reference writes at most #groups rows; N-row outputs are what a caller that does not know the group
        # count has to provide, but at 5e8+ rows they only cost memory: both arms get the same bound
        self.out_keys = torch.empty(min(self.N, 1 << 26), dtype=torch.int64, device="cuda")
        self.out_vals = torch.empty(min(self.N, 1 << 26), dtype=torch.int64, device="cuda")
        ffi, lib = api.ffi, api.lib
        self.ctx = ffi.new("gdf_context*")
        lib.gdf_context_view(self.ctx, 0, lib.GDF_HASH, 0, 0, 0)
        self.rows_per_step = self.N
        self.groups = None

    def step(self):
        ffi, lib = self.api.ffi, self.api.lib
        k, v = self.api.column(self.keys), self.api.column(self.vals)
        ok, ov = self.api.column(self.out_keys), self.api.column(self.out_vals)
        self.api.check(lib.gdf_group_by_sum(1, ffi.new("gdf_column*[]", [k]), v, ffi.NULL,
                                            ffi.new("gdf_column*[]", [ok]), ov, self.ctx), "gdf_group_by_sum")
        self.groups = int(ov.size)

    def algorithmic_bytes(self):  # 16 B/row in + 16 B/group out
        return 16 * self.N + 16 * self.groups

    def check(self):
        self.step()
        g = self.groups
        total_ok = int(self.out_vals[:g].sum().item()) == int(self.vals.sum().item())
        uniq = torch.unique(self.out_keys[:g]).numel() == g
        return total_ok and uniq and g == torch.unique(self.keys).numel()


class FilterWorkload(object):
    name = "C2 gdf_filter 1e9 int64 rows uniform [0,10), == 3 (10 % selectivity)"

    def __init__(self, api, scale):
        self.api = api
        self.N = int(1e9 * scale)
        self.col = torch.randint(0, 10, (self.N,), generator=gen(5), device="cuda", dtype=torch.int64)
        self.d_cols = torch.zeros(1, dtype=torch.int64, device="cuda")
        self.d_types = torch.zeros(1, dtype=torch.int32, device="cuda")
        self.val = torch.tensor([3], dtype=torch.int64, device="cuda")
        self.d_vals = torch.tensor([self.val.data_ptr()], dtype=torch.int64, device="cuda")
        self.d_indx = torch.empty(int(self.N * 0.11) + 1024, dtype=torch.int64, device="cuda")
        self.new_sz = api.ffi.new("size_t*")
        self.rows_per_step = self.N
        self.selected = None

    def step(self):
        ffi, lib = self.api.ffi, self.api.lib
        cols = ffi.new("gdf_column[]", 1)
        cols[0] = self.api.column(self.col)[0]
        self.api.check(lib.gdf_filter(self.N, cols, 1, ffi.cast("void**", self.d_cols.data_ptr()),
                                      ffi.cast("int*", self.d_types.data_ptr()), ffi.cast("void**", self.d_vals.data_ptr()),
                                      ffi.cast("size_t*", self.d_indx.data_ptr()), self.new_sz), "gdf_filter")
        self.selected = int(self.new_sz[0])

    def algorithmic_bytes(self):  # 8 B/row in + 8 B per selected index out
        return 8 * self.N + 8 * self.selected

    def check(self):
        self.step()
        k = self.selected
        idx = self.d_indx[:k]
        ok = k == int((self.col == 3).sum().item()) and bool((self.col[idx] == 3).all())
        return ok and bool((idx[1:] > idx[:-1]).all())


class StencilWorkload(object):
    name = "C2 variant: gpu_comparison_static_i64(== 3) + gpu_apply_stencil, 1e9 int64 rows, 10 % selectivity"

    def __init__(self, api, scale):
        self.api = api
        self.N = int(1e9 * scale)
        self.col = torch.randint(0, 10, (self.N,), generator=gen(5), device="cuda", dtype=torch.int64)
        self.stencil = torch.empty(self.N, dtype=torch.int8, device="cuda")
        self.svalid = torch.empty((self.N + 7) // 8, dtype=torch.uint8, device="cuda")
        self.out = torch.empty(self.N, dtype=torch.int64, device="cuda")      # the reference contract: output at input size
        self.ovalid = torch.empty((self.N + 7) // 8, dtype=torch.uint8, device="cuda")
        self.rows_per_step = self.N
        self.selected = None

    def _col(self, data, valid, code):
        ffi, lib = self.api.ffi, self.api.lib
        c = ffi.new("gdf_column*")
        lib.gdf_column_view(c, ffi.cast("void*", data.data_ptr()), ffi.cast("gdf_valid_type*", valid.data_ptr()) if valid is not None else ffi.NULL,
                            data.numel(), code)
        return c

    def step(self):
        lib = self.api.lib
        col = self._col(self.col, None, lib.GDF_INT64)
        st = self._col(self.stencil, self.svalid, lib.GDF_INT8)
        out = self._col(self.out, self.ovalid, lib.GDF_INT64)
        self.api.check(lib.gpu_comparison_static_i64(col, 3, st, lib.GDF_EQUALS), "gpu_comparison_static_i64")
        self.api.check(lib.gpu_apply_stencil(col, st, out), "gpu_apply_stencil")
        self.selected = int(out.size)

    def algorithmic_bytes(self):  # same figure as gdf_filter - the intermediate stencil traffic is not algorithmic
        return 8 * self.N + 8 * self.selected

    def check(self):
        self.step()
        k = self.selected
        return k == int((self.col == 3).sum().item()) and bool((self.out[:k] == 3).all())


class ReduceWorkload(object):
    def __init__(self, api, scale, masked):
        self.api, self.masked = api, masked
        self.N = int(1e9 * scale)
        self.name = "a15 gdf_sum_i64 over %.0e int64 rows%s" % (self.N, " with a 30 %-null validity mask" if masked else "")
        self.col = torch.randint(-1000, 1000, (self.N,), generator=gen(6), device="cuda", dtype=torch.int64)
        self.valid = None
        if masked:
            self.valid = torch.randint(0, 256, ((self.N + 7) // 8,), generator=gen(7), device="cuda", dtype=torch.uint8)
            self.valid &= torch.randint(0, 256, ((self.N + 7) // 8,), generator=gen(8), device="cuda", dtype=torch.uint8)
            self.valid |= torch.randint(0, 256, ((self.N + 7) // 8,), generator=gen(9), device="cuda", dtype=torch.uint8)  # P(bit) = 5/8
        self.scratch = torch.zeros(int(api.lib.gdf_reduce_optimal_output_size()), dtype=torch.int64, device="cuda")
        self.rows_per_step = self.N

    def step(self):
        ffi, lib = self.api.ffi, self.api.lib
        c = ffi.new("gdf_column*")
        lib.gdf_column_view(c, ffi.cast("void*", self.col.data_ptr()),
                            ffi.cast("gdf_valid_type*", self.valid.data_ptr()) if self.valid is not None else ffi.NULL, self.N, lib.GDF_INT64)
        self.api.check(lib.gdf_sum_i64(c, ffi.cast("int64_t*", self.scratch.data_ptr()), self.scratch.numel()), "gdf_sum_i64")

    def algorithmic_bytes(self):  # (8 B + 1 bit) per row
        return 8 * self.N + (self.N // 8 if self.masked else 0)

    def check(self):
        self.step()
        got = int(self.scratch[0].item())
        if not self.masked:
            return got == int(self.col.sum().item())
        want, chunk = 0, 1 << 27                                    # expand the bitmask in pieces (test-side only)
        shifts = torch.arange(8, device="cuda", dtype=torch.uint8)
        for lo in range(0, self.N, chunk):
            hi = min(self.N, lo + chunk)
            bits = ((self.valid[lo // 8:(hi + 7) // 8].unsqueeze(1) >> shifts) & 1).flatten()[: hi - lo].bool()
            want += int(self.col[lo:hi][bits].sum().item())
        return got == want


class BinaryAddWorkload(object):
    """a16: gdf_add_i64 on two int64 columns (C1 is the same call at 1M int32 rows: 12 MB, launch-latency bound)."""

    def __init__(self, api, scale, rows, dtype=torch.int64):
        self.api, self.N, self.dtype = api, max(int(rows * scale), 1024), dtype
        w = 8 if dtype == torch.int64 else 4
        self.name = "a16 gdf_add_%s over %.1e rows" % ("i64" if w == 8 else "i32", self.N)
        self.a = torch.randint(-10000, 10000, (self.N,), generator=gen(11), device="cuda", dtype=dtype)
        self.b = torch.randint(-10000, 10000, (self.N,), generator=gen(12), device="cuda", dtype=dtype)
        self.o = torch.empty_like(self.a)
        self.rows_per_step, self.w = self.N, w

    def step(self):
        lib = self.api.lib
        fn = lib.gdf_add_i64 if self.w == 8 else lib.gdf_add_i32
        self.api.check(fn(self.api.column(self.a), self.api.column(self.b), self.api.column(self.o)), "gdf_add")

    def algorithmic_bytes(self):
        return 3 * self.w * self.N

    def check(self):
        self.step()
        return bool((self.o == self.a + self.b).all())


class HashPartitionWorkload(object):
    """a17: gdf_hash_partition of (int64 key, int64 payload) rows into 8 partitions (MurmurHash3 row hash)."""

    def __init__(self, api, scale, rows=2.5e8, nparts=8):
        self.api, self.N, self.nparts = api, max(int(rows * scale), 1024), nparts
        self.name = "a17 gdf_hash_partition %.1e rows x (int64 key, int64 payload) -> %d partitions" % (self.N, nparts)
        self.k = torch.randint(0, 1 << 40, (self.N,), generator=gen(13), device="cuda", dtype=torch.int64)
        self.v = torch.arange(self.N, device="cuda", dtype=torch.int64)
        self.ok, self.ov = torch.empty_like(self.k), torch.empty_like(self.v)
        self.rows_per_step = self.N
        self.offsets = None

    def step(self):
        ffi, lib = self.api.ffi, self.api.lib
        self._ins = [self.api.column(self.k), self.api.column(self.v)]     # cffi arrays of pointers do not keep their targets alive
        ins = ffi.new("gdf_column*[]", self._ins)
        self._outs = [self.api.column(self.ok), self.api.column(self.ov)]
        outs = ffi.new("gdf_column*[]", self._outs)
        offs = ffi.new("int[]", self.nparts)
        self.api.check(lib.gdf_hash_partition(2, ins, ffi.new("int[]", [0]), 1, self.nparts, outs, offs, lib.GDF_HASH_MURMUR3),
                       "gdf_hash_partition")
        self.offsets = list(offs)

    def algorithmic_bytes(self):
        return 2 * 16 * self.N

    def check(self):
        """rows keep their (key, payload) pairing; partition sizes add up; the payload is a permutation."""
        self.step()
        ok = self.offsets[0] == 0 and all(a <= b for a, b in zip(self.offsets, self.offsets[1:] + [self.N]))
        ok = ok and bool((self.k[self.ov] == self.ok).all())
        return ok and int(self.ov.sum().item()) == self.N * (self.N - 1) // 2


class JoinGatherWorkload(object):
    """f1: gdf_inner_join with result_cols - [left payload, key, right payload] materialised by the fused gather."""

    def __init__(self, api, scale, P=2e8, B=2e7):
        self.api, self.P, self.B = api, max(int(P * scale), 4096), max(int(B * scale), 1024)
        self.name = "f1 gdf_inner_join %.0e x %.0e int64 + result_cols (int64 payload on both sides)" % (self.P, self.B)
        self.build = torch.randperm(self.B, generator=gen(14), device="cuda", dtype=torch.int64)
        self.probe = torch.randint(0, self.B, (self.P,), generator=gen(15), device="cuda", dtype=torch.int64)
        self.lp = torch.arange(self.P, device="cuda", dtype=torch.int64) * 3
        self.rp = torch.arange(self.B, device="cuda", dtype=torch.int64) * 7
        ffi, lib = api.ffi, api.lib
        self.ctx = ffi.new("gdf_context*")
        lib.gdf_context_view(self.ctx, 0, lib.GDF_HASH, 0, 0, 0)
        self.rows_per_step = self.P + self.B
        self.pairs = None

    def call(self):
        ffi, lib = self.api.ffi, self.api.lib
        self._cols = [self.api.column(t) for t in (self.lp, self.probe, self.build, self.rp)]
        la, ra = ffi.new("gdf_column*[]", self._cols[:2]), ffi.new("gdf_column*[]", self._cols[2:])
        self.res = [ffi.new("gdf_column*") for _ in range(3)]
        self.out_l, self.out_r = ffi.new("gdf_column*"), ffi.new("gdf_column*")
        self.api.check(lib.gdf_inner_join(la, 2, ffi.new("int[]", [1]), ra, 2, ffi.new("int[]", [0]), 1, 3,
                                          ffi.new("gdf_column*[]", self.res), self.out_l, self.out_r, self.ctx), "gdf_inner_join")
        self.pairs = int(self.out_l.size)

    def free(self):
        for c in self.res + [self.out_l, self.out_r]:
            if c.data != self.api.ffi.NULL:
                self.api.lib.gdf_column_free(c)

    def step(self):
        self.call()
        self.free()

    def algorithmic_bytes(self):   # the join's own 8 B/row in + 8 B/pair out, plus 3 materialised int64 columns
        return 8 * (self.P + self.B) + 8 * self.pairs + 3 * 8 * self.pairs

    def check(self):
        self.call()
        n, ffi = self.pairs, self.api.ffi
        li, ri = alias(ffi, self.out_l.data, n, np.int32).long(), alias(ffi, self.out_r.data, n, np.int32).long()
        a, k, b = (alias(ffi, c.data, n, np.int64) for c in self.res)
        ok = n == self.P and bool((a == self.lp[li]).all()) and bool((k == self.probe[li]).all()) and bool((b == self.rp[ri]).all())
        ok = ok and bool((self.probe[li] == self.build[ri]).all())
        del li, ri, a, k, b
        self.free()
        return ok


class C5Workload(object):
    name = "C5 gdf_left_join composite (int64,int32) key, 30 % null rows, 5e8 x 5e7"

    def __init__(self, api, scale):
        self.api = api
        self.NL, self.NR = int(5e8 * scale), int(5e7 * scale)

        def side(n, seed):
            g = gen(seed)
            k0 = torch.randint(0, self.NR, (n,), generator=g, device="cuda", dtype=torch.int64)
            k1 = torch.randint(0, 4, (n,), generator=g, device="cuda", dtype=torch.int32)
            null_row = torch.rand(n, generator=g, device="cuda") < 0.3
            which = torch.rand(n, generator=g, device="cuda") < 0.5
            return k0, k1, ~(null_row & which), ~(null_row & ~which)
        self.l0, self.l1, self.lv0, self.lv1 = side(self.NL, 100)
        self.r0, self.r1, self.rv0, self.rv1 = side(self.NR, 200)
        self.lm, self.rm = [pack_bits(self.lv0), pack_bits(self.lv1)], [pack_bits(self.rv0), pack_bits(self.rv1)]
        ffi, lib = api.ffi, api.lib
        self.ctx = ffi.new("gdf_context*")
        lib.gdf_context_view(self.ctx, 0, lib.GDF_HASH, 0, 0, 0)
        self.idx = ffi.new("int[]", [0, 1])
        self.out_l, self.out_r = ffi.new("gdf_column*"), ffi.new("gdf_column*")
        self.rows_per_step = self.NL + self.NR
        self.pairs = None

    def _cols(self, k0, k1, masks):
        ffi, lib = self.api.ffi, self.api.lib
        cols = []
        for t, m, code in ((k0, masks[0], lib.GDF_INT64), (k1, masks[1], lib.GDF_INT32)):
            c = ffi.new("gdf_column*")
            lib.gdf_column_view_augmented(c, ffi.cast("void*", t.data_ptr()), ffi.cast("gdf_valid_type*", m.data_ptr()), t.numel(), code, 1)
            cols.append(c)
        return cols, ffi.new("gdf_column*[]", cols)

    def call(self):
        ffi, lib = self.api.ffi, self.api.lib
        lk, la = self._cols(self.l0, self.l1, self.lm)
        rk, ra = self._cols(self.r0, self.r1, self.rm)
        self.api.check(lib.gdf_left_join(la, 2, self.idx, ra, 2, self.idx, 2, 0, ffi.NULL, self.out_l, self.out_r, self.ctx), "gdf_left_join")
        self.pairs = int(self.out_l.size)

    def free(self):
        for o in (self.out_l, self.out_r):
            if o.data != self.api.ffi.NULL:
                self.api.lib.gdf_column_free(o)

    def step(self):
        self.call()
        self.free()

    def algorithmic_bytes(self):  # (12 B + 2 bits) per input row + 8 B per output pair
        return (12 * (self.NL + self.NR) + (self.NL + self.NR) // 4) + 8 * self.pairs

    def check(self):
        """every left row appears; null left rows exactly once and with -1; matched pairs join equal valid keys."""
        self.call()
        n = self.pairs
        gl = alias(self.api.ffi, self.out_l.data, n, np.int32)
        gr = alias(self.api.ffi, self.out_r.data, n, np.int32)
        seen = torch.zeros(self.NL, dtype=torch.int32, device="cuda")
        seen.index_add_(0, gl.long(), torch.ones_like(gl))
        lvalid = self.lv0 & self.lv1
        ok = bool((seen >= 1).all()) and bool((seen[~lvalid] == 1).all())
        matched = gr >= 0
        ml, mr = gl[matched].long(), gr[matched].long()
        ok = ok and bool((self.l0[ml] == self.r0[mr]).all()) and bool((self.l1[ml] == self.r1[mr]).all())
        ok = ok and bool(lvalid[ml].all()) and bool((self.rv0 & self.rv1)[mr].all())
        unmatched_null = gl[~matched].long()
        del seen, gl, gr, ml, mr, unmatched_null
        self.free()
        return ok


def pack_bits(bits):
    """bool[n] -> Arrow LSB-first bitmask (synthetic-input preparation, not the product path)."""
    n = bits.numel()
    pad = (-n) % 8
    b = torch.cat([bits, torch.zeros(pad, dtype=torch.bool, device=bits.device)]).view(-1, 8).to(torch.uint8)
    w = torch.tensor([1, 2, 4, 8, 16, 32, 64, 128], dtype=torch.uint8, device=bits.device)
    return (b * w).sum(1).to(torch.uint8)


def run_workload(api, wl, args, peak_gbs, clocks):
    api.profile_begin()
    ms, (t0, t1) = timed_steps(wl.step, args.warmup, args.steps)
    prof = api.profile_end()
    # warm-up launches are included in prof: scale to the timed share by launches / (W+K)
    total_calls = args.warmup + args.steps
    kernels = {k: {"launches_per_step": v["launches"] / total_calls, "ms_per_step": v["ms"] / total_calls} for k, v in prof.items()}
    res = {
        "workload": wl.name, "ms_per_step": ms, "rows_per_step": wl.rows_per_step,
        "rows_per_s": wl.rows_per_step / (ms * 1e-3),
        "algorithmic_GB": wl.algorithmic_bytes() / 1e9,
        "path_GBps": wl.algorithmic_bytes() / 1e9 / (ms * 1e-3),
        "path_frac_of_measured_hbm": wl.algorithmic_bytes() / 1e9 / (ms * 1e-3) / peak_gbs,
        "kernels": kernels, "clocks": clocks.summarise(clocks.window(t0, t1)),
    }
    return res


# per-kernel algorithmic bytes per launch (compulsory reads + writes of THAT kernel), see DESIGN.md
def kernel_bytes(name, wl):
    """Algorithmic bytes of ONE launch of the kernel(s) timed under `name`: what that kernel must read and write."""
    if isinstance(wl, JoinWorkload):   # C3 takes the compact path: 8-byte {key32, tag32} pairs, 8-byte slots
        P, B, pairs = wl.P, wl.B, wl.pairs
        if name == "join_part_probe":
            return 8 * P + 8 * pairs             # pairs in, index pairs out
        if name == "join_part_scatter":
            return (8 + 8) * (P + B) / 2.0       # two launches per step (build side, probe side): average
        if name == "join_part_hist":
            return 8 * (P + B) / 2.0
        if name == "join_part_build":
            return 8 * B + 8 * 2 * B             # pairs in, slots (load factor 0.5) written
        return None
    if isinstance(wl, BinaryAddWorkload) and name == "binary_op":
        return wl.algorithmic_bytes()
    if isinstance(wl, HashPartitionWorkload) and name == "hash_partition":
        return wl.algorithmic_bytes()
    if isinstance(wl, JoinGatherWorkload):
        if name == "join_gather":
            return (4 + 8 + 8) * wl.pairs * 3 / 2.0   # two launches (left side: payload + key, right side: payload): index in, value in, value out
        if name == "join_part_probe":
            return 8 * wl.P + 8 * wl.pairs
    if isinstance(wl, GroupbyWorkload):
        if name == "groupby_build_fast":
            return 16 * wl.N
    if isinstance(wl, FilterWorkload):
        if name == "select":
            return 8 * wl.N + 8 * wl.selected
    if isinstance(wl, StencilWorkload):
        if name == "select":
            return 1 * wl.N + wl.N // 8 + 16 * wl.selected   # stencil bytes + mask in, selected values gathered and written
        if name == "compare_static":
            return 8 * wl.N + wl.N + wl.N // 8               # column in, int8 stencil + its mask out
    if isinstance(wl, ReduceWorkload):
        if name == "reduce":
            return wl.algorithmic_bytes()
    if isinstance(wl, C5Workload):
        n = wl.NL + wl.NR
        if name == "join_part_probe":
            return 16 * wl.NL + 8 * wl.pairs                 # {key 8, tag 4, key2 4} pairs in, index pairs out
        if name == "join_part_scatter":
            return (12 + 16) * n / 2.0 + n / 8.0             # two launches: keys + masks in, 16-byte {key,tag,key2} out
        if name == "join_part_count":
            return 16 * wl.NL
    return None


TRAFFIC_FILE = os.path.join(ROOT, "profiles", "r02_traffic_full_size.json")


def csrc_sha1():
    """Hash of the kernel sources: a traffic record only describes the kernels it was captured from."""
    import glob
    import hashlib
    h = hashlib.sha1()
    for path in sorted(glob.glob(os.path.join(ROOT, "libgdf_b200", "csrc", "*"))):
        if path.endswith((".cu", ".cuh", ".h", ".cpp")):
            h.update(os.path.basename(path).encode())
            h.update(open(path, "rb").read())
    return h.hexdigest()


def measured_traffic(workload_key, kernel, scale, launches_per_step=1.0):
    """DRAM bytes per launch (per-step bytes of everything timed under `kernel` / its launches per step) from the full-size
    ncu capture (tools/traffic_full.sh -> profiles/r02_traffic_full_size.json).  The file records the hash of csrc/ it was
    captured from; a record from other sources is stale and is NOT reported."""
    if scale != 1.0 or not os.path.isfile(TRAFFIC_FILE):
        return None
    data = json.load(open(TRAFFIC_FILE))
    if data.get("csrc_sha1") != csrc_sha1():
        return None
    rec = data.get("entries", {}).get(workload_key, {}).get(kernel)
    return int((rec["dram_read"] + rec["dram_write"]) / max(launches_per_step, 1e-9)) if rec else None


def roofline_for(res, wl, peak_gbs, peak_kind, scale=1.0, key=None):
    ks = res["kernels"]
    if not ks:
        return None
    top = max(ks, key=lambda k: ks[k]["ms_per_step"])
    per_launch_ms = ks[top]["ms_per_step"] / max(ks[top]["launches_per_step"], 1e-9)
    nbytes = kernel_bytes(top, wl)
    if nbytes is None:
        return {"kernel": top, "bound": "hbm", "achieved": None, "peak": peak_gbs, "unit": "GB/s", "frac": None, "traffic": None}
    achieved = nbytes / 1e9 / (per_launch_ms * 1e-3)
    return {"kernel": top, "bound": "hbm", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
            "frac": achieved / peak_gbs, "traffic": measured_traffic(key, top, scale, ks[top]["launches_per_step"]), "peak_source": peak_kind,
            "algorithmic_bytes_per_launch": nbytes, "avg_launch_ms": per_launch_ms,
            "share_of_step": ks[top]["ms_per_step"] / res["ms_per_step"]}


def cpu_baseline_join(scale):
    """C oracle port (oracle/gdf_oracle.c, single thread) on a bounded sample of C3: 5e7 x 5e6 (~10 s)."""
    import oracle
    rng = np.random.RandomState(SEED % (2 ** 32))
    B, P = max(int(5e6 * min(scale * 10, 1.0)), 1000), max(int(5e7 * min(scale * 10, 1.0)), 10000)
    build = rng.permutation(B).astype(np.int64)
    probe = rng.randint(0, B, P).astype(np.int64)
    t0 = time.perf_counter()
    l, r = oracle.join(oracle.JOIN_INNER, [probe], [build])
    dt = time.perf_counter() - t0
    assert len(l) == P
    return {"value": (P + B) / dt, "unit": "rows/s", "cores": 1, "kind": "port",
            "sample": "C3 shape at %d x %d rows (C oracle port, 1 thread, %.1f s)" % (P, B, dt)}


def load_peak():
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(peaks_path):
        return float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    return 6650.0, "B200_PROFILING.md fallback (of fallback)"


def bench_dist(args, rank, world, local_rank):
    """N > 1: C3 sharded over the ranks (strong scaling: the 1e9 x 1e8 tables are block-distributed),
    one hash-partition + all-to-all + local join per step; max over ranks of the device time."""
    import torch.distributed as dist
    from libgdf_b200 import dist as D
    if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", ""):
        os.environ["NCCL_DEBUG"] = "WARN"      # stdout carries exactly one JSON line
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    api = Api("b200")
    ops = D.GdfOps()
    peer = None
    if args.exchange == "p2p":
        if D.PeerExchange.available(ops):
            peer = D.PeerExchange(ops)
            peer.one_pass = not args.two_pass
            peer.overlap_build = not args.no_overlap
            peer.use_symm_mem = args.symm_mem
            if args.scatter_ctas >= 0:
                peer.scatter_ctas_per_sm = args.scatter_ctas
            if args.xjoin_rpp_log2:
                peer.rows_per_partition = 1 << args.xjoin_rpp_log2
        elif rank == 0:
            print("bench.py: CUDA IPC peer mapping unavailable on this box, falling back to --exchange nccl", file=sys.stderr)
    peak_gbs, peak_kind = load_peak()
    clocks = Clocks(local_rank)
    P, B = int(1e9 * args.scale), int(1e8 * args.scale)
    plo, phi = D.shard_bounds(P, world, rank)
    blo, bhi = D.shard_bounds(B, world, rank)
    perm = torch.randperm(B, generator=gen(1), device="cuda", dtype=torch.int64)       # same on every rank
    build = perm[blo:bhi].clone()
    probe = torch.randint(0, B, (phi - plo,), generator=gen(1000 * (rank + 1)), device="cuda", dtype=torch.int64)

    def all_max(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def all_sum(x):
        t = torch.tensor([x], dtype=torch.int64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return int(t.item())

    # ---- parity properties at full size: every probe row appears exactly once, every pair joins equal keys
    gl, gr = D.distributed_join("inner", probe, build, plo, blo, ops, peer=peer)
    pairs_local = gl.numel()
    pairs = all_sum(pairs_local)
    id_sum = all_sum(int(gl.long().sum().item()))
    full_probe = torch.empty(P, dtype=torch.int64, device="cuda")
    sizes = [D.shard_bounds(P, world, r) for r in range(world)]
    dist.all_gather([full_probe[lo:hi] for lo, hi in sizes], probe)
    keys_ok = bool((full_probe[gl.long()] == perm[gr.long()]).all())
    del full_probe
    parity_ok = all_sum(int(keys_ok)) == world and pairs == P and id_sum == P * (P - 1) // 2
    del gl, gr, perm
    torch.cuda.empty_cache()

    timings = {}

    def step():
        a, b = D.distributed_join("inner", probe, build, plo, blo, ops, timings=timings, peer=peer, global_rows=(P, B))
        del a, b

    def timed(fn, warmup, steps):
        for _ in range(warmup):
            fn()
        timings.clear()
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        dist.barrier()
        t1 = time.perf_counter()
        return all_max(e0.elapsed_time(e1) / steps), (t0, t1)

    api.lib.gdfx_profile_enable(0)
    ms, (t0, t1) = timed(step, args.warmup, args.steps)
    phases = {k: v / args.steps for k, v in D.resolve_timings(timings).items()}
    # second, short pass with the per-kernel event timers on (kept out of the headline timing)
    api.profile_begin()
    step()
    torch.cuda.synchronize()
    prof = api.profile_end()
    api.lib.gdfx_profile_enable(0)
    kernels = {k: {"launches_per_step": v["launches"], "ms_per_step": v["ms"]} for k, v in prof.items()}

    # ---- e2e: pinned host shards -> H2D -> distributed join -> D2H of this rank's result
    e2e = None
    if not args.no_e2e:
        h_probe = torch.empty(probe.numel(), dtype=torch.int64, pin_memory=True).copy_(probe)
        h_build = torch.empty(build.numel(), dtype=torch.int64, pin_memory=True).copy_(build)
        d_probe, d_build = torch.empty_like(probe), torch.empty_like(build)
        cap = int(pairs_local * 1.5) + 1024
        h_l = torch.empty(cap, dtype=torch.int32, pin_memory=True)
        h_r = torch.empty(cap, dtype=torch.int32, pin_memory=True)

        def e2e_step():
            d_probe.copy_(h_probe, non_blocking=True)
            d_build.copy_(h_build, non_blocking=True)
            a, b = D.distributed_join("inner", d_probe, d_build, plo, blo, ops, peer=peer, global_rows=(P, B))
            h_l[: a.numel()].copy_(a, non_blocking=True)
            h_r[: b.numel()].copy_(b, non_blocking=True)
            torch.cuda.synchronize()
        e_steps = max(1, min(args.steps, 3))
        e_ms, _ = timed(e2e_step, 1, e_steps)
        e2e = {"value": (P + B) / (e_ms * 1e-3), "unit": "rows/s", "h2d_bytes_per_step": 8 * (P + B),
               "d2h_bytes_per_step": 8 * pairs, "ms_per_step": e_ms, "steps": e_steps,
               "note": "bytes are the whole job's (all ranks); each rank copies its own shard"}
    # ---- the other sharded configs of BASELINE.json under the same launch: C4 group-by-sum and C5 left join ----
    del probe, build
    torch.cuda.empty_cache()
    workloads = {}
    short_steps, short_warm = max(1, min(args.steps, 5)), max(1, min(args.warmup, 2))

    def run_sharded(step_fn, steps, warm):
        tm = {}

        def one():
            r = step_fn(tm)
            del r
        for _ in range(warm):
            one()
        tm.clear()
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            one()
        e1.record()
        torch.cuda.synchronize()
        dist.barrier()
        return all_max(e0.elapsed_time(e1) / steps), {k: v / steps for k, v in D.resolve_timings(tm).items()}

    if not args.only or "groupby" in args.only:
        try:
            N4, G4 = int(1e9 * args.scale), max(int(1e6 * args.scale), 16)
            lo4, hi4 = D.shard_bounds(N4, world, rank)
            ranks_ = torch.arange(1, G4 + 1, device="cuda", dtype=torch.float64)
            cdf = torch.cumsum(ranks_.pow(-1.05), 0)
            cdf /= cdf[-1].clone()
            ids = torch.randperm(G4, generator=gen(3), device="cuda", dtype=torch.int64) * 7919 + 13   # same on every rank
            keys = torch.empty(hi4 - lo4, dtype=torch.int64, device="cuda")
            g4 = gen(2 + 1000 * rank)
            for a in range(0, hi4 - lo4, 1 << 27):
                b = min(hi4 - lo4, a + (1 << 27))
                u = torch.rand(b - a, generator=g4, device="cuda", dtype=torch.float64)
                keys[a:b] = ids[torch.searchsorted(cdf, u).clamp_(max=G4 - 1)]
                del u
            vals = torch.randint(0, 1000, (hi4 - lo4,), generator=gen(4 + 1000 * rank), device="cuda", dtype=torch.int64)
            gk, gv = D.distributed_group_by_sum(keys, vals, ops)
            # parity at full size: sum of sums = sum of values; every key on exactly one rank; #groups = #distinct keys
            tot_ok = all_sum(int(gv.sum().item())) == all_sum(int(vals.sum().item()))
            mine = torch.zeros(G4, dtype=torch.int32, device="cuda")
            pos = torch.searchsorted(torch.sort(ids).values, gk)
            mine.index_add_(0, pos, torch.ones_like(pos, dtype=torch.int32))
            dist.all_reduce(mine)
            present = torch.zeros(G4, dtype=torch.int32, device="cuda")
            present.index_fill_(0, torch.searchsorted(torch.sort(ids).values, torch.unique(keys)), 1)
            dist.all_reduce(present, op=dist.ReduceOp.MAX)
            owner_ok = bool((mine == present).all())
            groups = all_sum(gk.numel())
            del gk, gv, mine, present, pos
            ms4, ph4 = run_sharded(lambda tm: D.distributed_group_by_sum(keys, vals, ops, timings=tm), args.steps, args.warmup)
            workloads["groupby"] = {"workload": GroupbyWorkload.name + ", rows block-distributed over %d ranks" % world,
                                    "ms_per_step": ms4, "rows_per_step": N4, "rows_per_s": N4 / (ms4 * 1e-3), "groups": groups,
                                    "parity_properties_ok": bool(tot_ok and owner_ok), "phases_ms_rank0": ph4,
                                    "steps": args.steps, "warmup": args.warmup,
                                    "exchange": "local gdf_group_by_sum, gdf_hash_partition of the partials, one NCCL all_to_all per "
                                                "column (<= 16 MB per rank), merging gdf_group_by_sum"}
            del keys, vals
        except Exception as exc:
            workloads["groupby"] = {"error": str(exc)[:300]}
        torch.cuda.empty_cache()
    if not args.only or "c5" in args.only:
        try:
            NL, NR = int(5e8 * args.scale), int(5e7 * args.scale)
            llo, lhi = D.shard_bounds(NL, world, rank)
            rlo, rhi = D.shard_bounds(NR, world, rank)

            def side(n, seed):
                g = gen(seed)
                k0 = torch.randint(0, NR, (n,), generator=g, device="cuda", dtype=torch.int64)
                k1 = torch.randint(0, 4, (n,), generator=g, device="cuda", dtype=torch.int32)
                null_row = torch.rand(n, generator=g, device="cuda") < 0.3
                which = torch.rand(n, generator=g, device="cuda") < 0.5
                return k0, k1, ~(null_row & which), ~(null_row & ~which)
            l0, l1, lv0, lv1 = side(lhi - llo, 100 + rank)
            r0, r1, rv0, rv1 = side(rhi - rlo, 200 + rank)
            lmask, rmask = [pack_bits(lv0), pack_bits(lv1)], [pack_bits(rv0), pack_bits(rv1)]
            gl, gr = D.distributed_left_join_masked([l0, l1], lmask, [r0, r1], rmask, llo, rlo, ops, peer=peer)
            seen = torch.zeros(NL, dtype=torch.int32, device="cuda")
            seen.index_add_(0, gl.long(), torch.ones_like(gl))
            dist.all_reduce(seen)
            lvalid = lv0 & lv1
            ok5 = bool((seen[llo:lhi] >= 1).all()) and bool((seen[llo:lhi][~lvalid] == 1).all())
            del seen

            def gathered(x, total):
                full = torch.empty(total, dtype=x.dtype, device="cuda")
                dist.all_gather([full[a:b] for a, b in (D.shard_bounds(total, world, k) for k in range(world))], x)
                return full
            L0, L1, LV = gathered(l0, NL), gathered(l1, NL), gathered(lvalid, NL)
            R0, R1, RV = gathered(r0, NR), gathered(r1, NR), gathered(rv0 & rv1, NR)
            m = gr >= 0
            ml, mr = gl[m].long(), gr[m].long()
            ok5 = ok5 and bool((L0[ml] == R0[mr]).all()) and bool((L1[ml] == R1[mr]).all()) and bool(LV[ml].all()) and bool(RV[mr].all())
            pairs5 = all_sum(gl.numel())
            ok5 = all_sum(int(ok5)) == world
            del L0, L1, LV, R0, R1, RV, gl, gr, ml, mr, m
            torch.cuda.empty_cache()
            ms5, ph5 = run_sharded(lambda tm: D.distributed_left_join_masked([l0, l1], lmask, [r0, r1], rmask, llo, rlo, ops,
                                                                            timings=tm, peer=peer), short_steps, short_warm)
            workloads["c5"] = {"workload": C5Workload.name + ", rows block-distributed over %d ranks" % world, "ms_per_step": ms5,
                               "rows_per_step": NL + NR, "rows_per_s": (NL + NR) / (ms5 * 1e-3), "output_pairs": pairs5,
                               "parity_properties_ok": ok5, "phases_ms_rank0": ph5, "steps": short_steps, "warmup": short_warm}
            del l0, l1, r0, r1
        except Exception as exc:
            workloads["c5"] = {"error": str(exc)[:300]}
        torch.cuda.empty_cache()
    clocks.stop()
    if rank == 0:
        top = max(kernels, key=lambda k: kernels[k]["ms_per_step"]) if kernels else None
        out = {"metric": "rows_per_sec_hash_inner_join_int64", "value": (P + B) / (ms * 1e-3), "unit": "rows/s",
               "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
               "scaling": "strong", "vs_baseline": None, "dtype": "int64",
               "data": "synthetic (seeded torch device RNG, SURVEY.md 8d)", "impl": "b200",
               "config": {"workload": JoinWorkload.name + ", rows block-distributed over %d ranks" % world,
                          "probe_rows": P, "build_rows": B, "key_dtype": "int64", "rows_counted": "probe+build",
                          "parallelism": ("partition kernel storing into NVLink peer memory + local join (dp%d)" if peer else "hash-partition + NCCL all_to_all + local join (dp%d)") % world,
                          "l2_policy": "per-rank inputs (%.1f GB) larger than L2, no flush" % (8 * (P + B) / world / 1e9)},
               "parity_properties_ok": parity_ok, "output_pairs": pairs, "phases_ms_rank0": phases,
               "kernels_rank0": kernels, "clocks": clocks.summarise(clocks.window(t0, t1)),
               "gpu_launches": int(sum(k["launches_per_step"] for k in kernels.values()) * args.steps) if kernels else None}
        if top:
            # dominant kernel of rank 0 on its share of the rows (P/N probe pairs, B/N build pairs)
            per_launch = {"join_part_probe": 12 * P / world + 8 * pairs / world,
                          "join_part_scatter": (8 + 12) * (P + B) / world / 4.0,   # 4 launches: exchange + local, 2 sides
                          "join_part_hist": 8 * (P + B) / world / 4.0,
                          "join_part_build": (12 + 16) * B / world}.get(top)
            ms_l = kernels[top]["ms_per_step"] / max(kernels[top]["launches_per_step"], 1)
            ach = per_launch / 1e9 / (ms_l * 1e-3) if per_launch else None
            out["roofline"] = {"kernel": top, "bound": "hbm", "achieved": ach, "peak": peak_gbs, "unit": "GB/s",
                               "frac": ach / peak_gbs if ach else None, "traffic": None, "peak_source": peak_kind,
                               "algorithmic_bytes_per_launch": per_launch, "avg_launch_ms": ms_l,
                               "note": "rank 0's dominant kernel on its 1/N share; the step as a whole is bounded by the "
                                       "NVLink all-to-all (phases_ms_rank0), %.2f GB sent per rank per step"
                                       % (12 * (P + B) / world * (world - 1) / world / 1e9)}
        if e2e:
            out["e2e"] = e2e
        out["workloads"] = workloads
        emit(out)
    if peer:
        peer.close()
    dist.destroy_process_group()
    return 0


def cpu_baseline_pyarrow(scale):
    """pyarrow CPU pass over the same kind of Arrow columns on all host cores, bounded samples
    (join 2e7 x 2e6, group-by 2e7 rows / 2e4 Zipf groups, filter 2e7 rows).  Reported, not a target."""
    import pyarrow as pa
    import pyarrow.compute as pc
    pa.set_cpu_count(os.cpu_count() or 1)
    rng = np.random.RandomState(SEED % (2 ** 32))
    f = min(scale * 10, 1.0)
    B, P = max(int(2e6 * f), 1000), max(int(2e7 * f), 10000)
    out = {"cores": pa.cpu_count(), "unit": "rows/s", "pyarrow": pa.__version__}
    build, probe = rng.permutation(B).astype(np.int64), rng.randint(0, B, P).astype(np.int64)
    lt = pa.table({"k": probe, "l": np.arange(P, dtype=np.int32)})
    rt = pa.table({"k": build, "r": np.arange(B, dtype=np.int32)})
    t0 = time.perf_counter()
    j = lt.join(rt, keys="k", join_type="inner")
    dt = time.perf_counter() - t0
    assert j.num_rows == P
    out["join"] = {"value": (P + B) / dt, "sample": "Table.join inner, %d x %d int64 keys, %.2f s" % (P, B, dt)}
    G = max(int(2e4 * f), 16)
    w = np.arange(1, G + 1, dtype=np.float64) ** -1.05
    keys = (rng.choice(G, size=P, p=w / w.sum()).astype(np.int64)) * 7919 + 13
    vals = rng.randint(0, 1000, P).astype(np.int64)
    t0 = time.perf_counter()
    g = pa.table({"k": keys, "v": vals}).group_by("k").aggregate([("v", "sum")])
    dt = time.perf_counter() - t0
    assert g.num_rows <= G
    out["groupby"] = {"value": P / dt, "sample": "group_by().aggregate(sum), %d rows, %d Zipf groups, %.2f s" % (P, G, dt)}
    col = pa.array(rng.randint(0, 10, P).astype(np.int64))
    t0 = time.perf_counter()
    idx = pc.indices_nonzero(pc.equal(col, 3))
    dt = time.perf_counter() - t0
    out["filter"] = {"value": P / dt, "sample": "indices_nonzero(equal(col, 3)), %d rows, %d selected, %.3f s" % (P, len(idx), dt)}
    return out


_REAL_STDOUT = None


def emit(obj):
    """The ONE JSON line goes to the process's original stdout; everything else any library prints to fd 1
    (e.g. NCCL's "NCCL version ..." banner) has been redirected to stderr by main()."""
    _REAL_STDOUT.write(json.dumps(obj) + "\n")
    _REAL_STDOUT.flush()


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    args = parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference" and rank != 0:
        return 0
    torch.cuda.set_device(local_rank)
    if world > 1 and args.impl == "b200":
        return bench_dist(args, rank, world, local_rank)

    peak_gbs, peak_kind = load_peak()
    ref_sample = None
    if args.impl == "reference":
        # Same config as the b200 arm: C3 at 1e9 x 1e8 and C2 at 1e9 rows through the reference's own kernels.
        # Only its group-by is bounded (see the C4 block below): 5e8 rows with uniform keys (int overflow above
        # 2^29 rows) plus a 5e6-row Zipf sample (its CAS loop serialises on the hot key).
        ref_sample = ("C3 1e9 x 1e8 and C2 1e9 rows at full size; C4 bounded: 5e8 rows uniform keys + 5e6-row Zipf sample"
                      if args.scale == 1.0 else "scale %g" % args.scale)

    try:
        api = Api(args.impl, args.lab)
    except Exception as exc:  # reference library not built
        emit({"impl": args.impl, "unavailable": str(exc)})
        return 0
    only = set(filter(None, args.only.split(",")))
    clocks = Clocks(local_rank)
    out = {"metric": "rows_per_sec_hash_inner_join_int64", "unit": "rows/s", "n_gpus": 1, "steps": args.steps,
           "warmup": args.warmup, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int64",
           "data": "synthetic (seeded torch device RNG, SURVEY.md 8d)", "impl": args.impl}
    workloads = {}

    # ---- headline: C3 inner join ----
    if not only or "join" in only:
        wl = JoinWorkload(api, args.scale)
        parity_ok = wl.check()
        res = run_workload(api, wl, args, peak_gbs, clocks)
        res["parity_properties_ok"] = parity_ok
        res["output_pairs"] = wl.pairs
        out.update({"value": res["rows_per_s"], "ms_per_step": res["ms_per_step"], "clocks": res["clocks"],
                    "config": {"workload": wl.name, "probe_rows": wl.P, "build_rows": wl.B, "key_dtype": "int64",
                               "rows_counted": "probe+build", "l2_policy": "inputs (8.8 GB) larger than L2, no flush",
                               "rmm": "pool"}})
        roof = roofline_for(res, wl, peak_gbs, peak_kind, args.scale, "join")
        if roof:
            out["roofline"] = roof
        out["gpu_launches"] = int(round(sum(k["launches_per_step"] for k in res["kernels"].values()) * args.steps)) if res["kernels"] else None
        if not args.no_e2e and args.impl == "b200":
            try:
                h2d, d2h = wl.e2e_setup()
                e_steps = max(1, min(args.steps, 3))
                e_ms, _ = timed_steps(wl.e2e_step, 1, e_steps)
                out["e2e"] = {"value": wl.rows_per_step / (e_ms * 1e-3), "unit": "rows/s", "h2d_bytes_per_step": h2d,
                              "d2h_bytes_per_step": d2h, "ms_per_step": e_ms, "steps": e_steps, "mode": "one call",
                              "result_ok": wl.e2e_verify()}
                try:   # beside it: streamed through the same C-ABI call in 8 probe pieces, copies overlap the joins
                    p_ms, _ = timed_steps(wl.e2e_step_pipelined, 1, e_steps)
                    out["e2e"]["pipelined"] = {"value": wl.rows_per_step / (p_ms * 1e-3), "ms_per_step": p_ms,
                                               "result_ok": wl.e2e_verify(),
                                               "mode": "probe column streamed in 8 pieces, one gdf_inner_join per piece, "
                                                       "H2D / join / D2H overlapped on three streams"}
                except Exception as exc:
                    out["e2e"]["pipelined_error"] = str(exc)[:200]
                for a in ("h_probe", "h_build", "d_probe", "d_build", "h_out_l", "h_out_r"):
                    delattr(wl, a)
                torch.cuda.synchronize()
            except Exception as exc:
                out["e2e"] = {"value": None, "unit": "rows/s", "error": str(exc)[:200]}
        workloads["join"] = res
        del wl
        torch.cuda.empty_cache()

    # ---- C4 group-by sum ----
    if not only or any(k.startswith("groupby") for k in only):
        import copy
        # (name, kwargs, args override): the b200 arm runs C4 itself (Zipf) and the uniform-key variant at 1e9 rows;
        # the reference arm is bounded: its table is 2N x 16 B and its int grid-stride arithmetic is unsafe above 2^29
        # rows, and under Zipf skew its CAS loop on the hot key's slot serialises (5 s per call at
        # 5e6 rows on this B200), so: uniform keys at 5e8 rows + a 5e6-row Zipf sample with at most 3 timed steps.
        short = copy.copy(args)
        short.steps, short.warmup = max(1, min(args.steps, 3)), max(1, min(args.warmup, 1))
        if args.impl == "reference":
            plan = [("groupby", dict(rows=5e6, zipf=True, groups=1e5), short),
                    ("groupby_uniform", dict(rows=5e8, zipf=False), short)]
        else:
            plan = [("groupby", dict(), args), ("groupby_uniform", dict(zipf=False), args),
                    ("groupby_uniform_5e8", dict(rows=5e8, zipf=False), args)]
        for key, kw, a in plan:
            if only and key not in only and "groupby_all" not in only:
                continue
            try:
                wl = GroupbyWorkload(api, args.scale, **kw)
                ok = wl.check()
                res = run_workload(api, wl, a, peak_gbs, clocks)
                res["parity_properties_ok"] = ok
                res["groups"] = wl.groups
                res["steps"], res["warmup"] = a.steps, a.warmup
                res["roofline"] = roofline_for(res, wl, peak_gbs, peak_kind, args.scale, key)
                workloads[key] = res
                del wl
            except Exception as exc:
                workloads[key] = {"error": str(exc)[:300]}
            torch.cuda.empty_cache()

    # ---- C2 filter ----
    if not only or "filter" in only:
        try:
            wl = FilterWorkload(api, args.scale)
            ok = wl.check()
            res = run_workload(api, wl, args, peak_gbs, clocks)
            res["parity_properties_ok"] = ok
            res["selected"] = wl.selected
            res["roofline"] = roofline_for(res, wl, peak_gbs, peak_kind, args.scale, "filter")
            workloads["filter"] = res
            del wl
        except Exception as exc:
            workloads["filter"] = {"error": str(exc)[:300]}
        torch.cuda.empty_cache()
    # ---- more rows of SURVEY 8a / 8d, each with its own roofline (b200 arm only) ----
    if args.impl == "b200":
        extra = [("filter_stencil", lambda: StencilWorkload(api, args.scale), lambda wl: {"selected": wl.selected}),
                 ("reduce_sum", lambda: ReduceWorkload(api, args.scale, False), lambda wl: {}),
                 ("reduce_sum_masked", lambda: ReduceWorkload(api, args.scale, True), lambda wl: {}),
                 ("add_c1", lambda: BinaryAddWorkload(api, 1.0, 1e6, torch.int32), lambda wl: {}),
                 ("add_i64_2.5e8", lambda: BinaryAddWorkload(api, args.scale, 2.5e8), lambda wl: {}),
                 ("hash_partition", lambda: HashPartitionWorkload(api, args.scale), lambda wl: {}),
                 ("join_result_cols", lambda: JoinGatherWorkload(api, args.scale), lambda wl: {"output_pairs": wl.pairs}),
                 ("c5", lambda: C5Workload(api, args.scale), lambda wl: {"output_pairs": wl.pairs})]
        for key, make, more in extra:
            if only and key not in only:
                continue
            try:
                wl = make()
                ok = wl.check()
                res = run_workload(api, wl, args, peak_gbs, clocks)
                res["parity_properties_ok"] = ok
                res.update(more(wl))
                res["roofline"] = roofline_for(res, wl, peak_gbs, peak_kind, args.scale, key)
                workloads[key] = res
                del wl
            except Exception as exc:
                workloads[key] = {"error": str(exc)[:300]}
            torch.cuda.empty_cache()
    clocks.stop()

    out["workloads"] = workloads
    if "value" not in out and workloads:   # debug runs with --only
        first = next(iter(workloads.values()))
        out.update({"value": first.get("rows_per_s"), "ms_per_step": first.get("ms_per_step"),
                    "config": {"workload": first.get("workload")}, "clocks": first.get("clocks")})
    if args.impl == "reference":
        out["cpu_baseline"] = {"value": out.get("value"), "unit": "rows/s", "cores": 1, "kind": "reference",
                               "sample": (ref_sample or "scale %g" % args.scale) + "; the reference is a CUDA library with "
                               "no CPU path: this arm runs its own kernels rebuilt for sm_100a (oracle/_ref) on the GPU"}
        out["e2e"] = {"value": out.get("value"), "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    elif not args.no_cpu:
        try:
            out["cpu_baseline"] = cpu_baseline_join(args.scale)
        except Exception as exc:
            out["cpu_baseline"] = {"value": None, "unit": "rows/s", "cores": 1, "kind": "port", "sample": "failed: %s" % exc}
        try:
            out["cpu_baseline_pyarrow"] = cpu_baseline_pyarrow(args.scale)
        except Exception as exc:
            out["cpu_baseline_pyarrow"] = {"error": str(exc)[:200]}
    emit(out)
    return 0


if __name__ == "__main__":
    sys.exit(main())
