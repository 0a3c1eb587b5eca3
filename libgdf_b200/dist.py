"""Multi-GPU layer: hash joins and hash group-bys sharded over the ranks of one torch.distributed
process group (one process per GPU, NCCL over NVLink 5 / NVSwitch).

The reference has no multi-device code at all (no NCCL/MPI call site, the join
even hard-codes ``prefetch(0)``, ref src/join/hash/join_compute_api.h:397).  Its only partitioner is the
single-GPU ``gdf_hash_partition`` (ref src/hashing.cu:559-654); this module is the layer one would put
on top of it, built from the same C-ABI operators:

    join      rows are block-distributed.  Every rank radix-partitions its local key column into
              G = world_size destination ranges of {key, global row id} pairs (``gdfx_partition_pairs``:
              the semantics of ``gdf_hash_partition`` - destination = fixed hash of the key - run by the
              join's own histogram + write-combining scatter kernels), exchanges the range sizes and then
              the ranges with ONE ``all_to_all_single`` per column, and joins what it received with
              ``gdfx_join_pairs`` = the single-GPU ``gdf_inner_join`` / ``gdf_left_join`` kernels with the
              travelling row ids as row tags, so the outputs are global row ids directly.  Equal keys
              always meet on one rank; output stays sharded.
    group-by  two-phase: local ``gdf_group_by_sum`` first (at most #groups partial rows per rank,
              which also removes Zipf skew: a hot key becomes ONE partial row per rank), then the
              partials are hash-partitioned, exchanged and merged with a second ``gdf_group_by_sum``.
    filter / reductions / binary ops   embarrassingly parallel: callers run the single-GPU operator per
              shard (global row index = local index + shard offset); nothing to exchange.

The exchange plan (``exchange``) is backend-agnostic torch.distributed code and is unit-tested on CPU
with the gloo backend and world_size 2 (tests/test_dist_cpu.py), where the per-shard operators are
injected by the test.  The product operators (``GdfOps``) are the CUDA C ABI and nothing else: there
is no CPU fallback in this module.
"""
import numpy as np
import torch
import torch.distributed as dist


# ------------------------------------------------------------------------------------------------
# per-shard operators: the gdf_* C ABI on CUDA tensors
# ------------------------------------------------------------------------------------------------
class GdfOps(object):
    """Single-GPU operators used by the distributed algorithms, all through libgdf.so."""

    def __init__(self):
        from . import columns as C
        from .libgdf_cffi import ffi, libgdf
        self.C, self.ffi, self.lib = C, ffi, libgdf
        self.ctx = ffi.new("gdf_context*")
        libgdf.gdf_context_view(self.ctx, 0, libgdf.GDF_HASH, 0, 0, 0)

    def hash_partition(self, cols, nparts):
        """cols[0] is the key.  Returns (partition-contiguous columns, start offset of each partition)."""
        C, ffi, lib = self.C, self.ffi, self.lib
        n = cols[0].numel()
        if n == 0:
            return [c.clone() for c in cols], [0] * nparts
        ins = [C.Column(c) for c in cols]
        outs = [C.Column(torch.empty_like(c)) for c in cols]
        offsets = ffi.new("int[]", nparts)
        lib.gdf_hash_partition(len(ins), C.column_array(ins), ffi.new("int[]", [0]), 1, nparts, C.column_array(outs),
                               offsets, lib.GDF_HASH_MURMUR3)
        return [o.data for o in outs], list(offsets)

    def partition_pairs(self, keys, id_base, nparts):
        """Split a key column into nparts destination ranges of {key, global row id = id_base + position}.
        Returns (keys_out, ids_out, start offsets).  One histogram + one write-combining scatter pass
        (gdfx_partition_pairs: the join's own radix-partition kernels with a destination hash)."""
        C, ffi, lib = self.C, self.ffi, self.lib
        n = keys.numel()
        out_k = torch.empty_like(keys)
        out_i = torch.empty(n, dtype=torch.int32, device=keys.device)
        offs = ffi.new("unsigned long long[]", nparts + 1)
        lib.gdfx_partition_pairs(C.Column(keys).cdata, id_base, nparts, ffi.cast("void*", out_k.data_ptr()),
                                 ffi.cast("int32_t*", out_i.data_ptr()), offs)
        return out_k, out_i, [int(offs[p]) for p in range(nparts)]

    def hash_partition_rows(self, cols, num_keys, nparts):
        """gdf_hash_partition of all `cols`, hashing the first `num_keys` of them (MurmurHash3 row hash)."""
        C, ffi, lib = self.C, self.ffi, self.lib
        n = cols[0].numel()
        if n == 0:
            return [c.clone() for c in cols], [0] * nparts
        ins = [C.Column(c) for c in cols]
        outs = [C.Column(torch.empty_like(c)) for c in cols]
        offsets = ffi.new("int[]", nparts)
        lib.gdf_hash_partition(len(ins), C.column_array(ins), ffi.new("int[]", list(range(num_keys))), num_keys, nparts,
                               C.column_array(outs), offsets, lib.GDF_HASH_MURMUR3)
        return [o.data for o in outs], list(offsets)

    def rows_valid_bytes(self, key_cols, valids):
        """int8 column: 1 where every key column's validity bit is set (valids[i] = packed mask tensor or None)."""
        C, ffi, lib = self.C, self.ffi, self.lib
        n = key_cols[0].numel()
        out = torch.empty(n, dtype=torch.int8, device=key_cols[0].device)
        # null_count is not used by the kernels; passing it avoids Column()'s host-side popcount of the mask
        cols = [C.Column(k, v, null_count=0 if v is None else 1) for k, v in zip(key_cols, valids)]
        lib.gdfx_rows_valid_to_bytes(C.column_array(cols), len(cols), ffi.cast("int8_t*", out.data_ptr()))
        return out

    def left_join_masked(self, lkeys, lok, rkeys, rok, lids, rids):
        """gdf_left_join on N key columns whose row validity arrives as byte columns (lok / rok); outputs are
        the travelling global row ids (-1 = no partner)."""
        C, ffi, lib = self.C, self.ffi, self.lib

        def with_mask(keys, ok):
            n = keys[0].numel()
            mask = torch.empty((n + 7) // 8, dtype=torch.uint8, device=keys[0].device)
            lib.gdfx_bytes_to_valid(ffi.cast("int8_t*", ok.data_ptr()), n, ffi.cast("gdf_valid_type*", mask.data_ptr()))
            return [C.Column(keys[0], mask, null_count=1)] + [C.Column(k) for k in keys[1:]]   # row-valid = AND over columns

        L, R = with_mask(lkeys, lok), with_mask(rkeys, rok)
        out_l, out_r = ffi.new("gdf_column*"), ffi.new("gdf_column*")
        idx = ffi.new("int[]", list(range(len(L))))
        lib.gdf_left_join(C.column_array(L), len(L), idx, C.column_array(R), len(R), idx, len(L), 0, ffi.NULL, out_l, out_r,
                          self.ctx)
        if int(out_l.size):
            lib.gdfx_remap_indices(out_l, ffi.cast("int32_t*", lids.data_ptr()), lids.numel())
            lib.gdfx_remap_indices(out_r, ffi.cast("int32_t*", rids.data_ptr()), rids.numel())
        return C.library_owned_to_torch(out_l), C.library_owned_to_torch(out_r)

    def partition_count(self, keys, nparts):
        ffi, lib = self.ffi, self.lib
        counts = ffi.new("unsigned long long[]", nparts)
        lib.gdfx_partition_count(self.C.Column(keys).cdata, nparts, counts)
        return [int(counts[p]) for p in range(nparts)]

    def partition_scatter_peer(self, keys, id_base, dst_key_ptrs, dst_id_ptrs, dst_offsets):
        """One pass: every {key, id} pair is stored straight into its destination's buffers (device
        addresses valid on this GPU: local memory or IPC-mapped peer memory)."""
        ffi, lib = self.ffi, self.lib
        n = len(dst_key_ptrs)
        kp = ffi.new("void*[]", [ffi.cast("void*", p) for p in dst_key_ptrs])
        ip = ffi.new("int32_t*[]", [ffi.cast("int32_t*", p) for p in dst_id_ptrs])
        off = ffi.new("unsigned long long[]", [int(o) for o in dst_offsets])
        lib.gdfx_partition_scatter_peer(self.C.Column(keys).cdata, id_base, n, kp, ip, off)

    def xjoin_count(self, keys, ranks, nlocal):
        """Rows of `keys` per (destination rank, receiver-local partition) + OR of the keys' high words."""
        ffi, lib = self.ffi, self.lib
        counts = ffi.new("unsigned long long[]", ranks * nlocal)
        hi = ffi.new("unsigned*")
        lib.gdfx_xjoin_count(self.C.Column(keys).cdata, ranks, nlocal, counts, hi)
        return [int(counts[i]) for i in range(ranks * nlocal)], int(hi[0])

    def xjoin_scatter(self, keys, id_base, ranks, nlocal, dst_ptrs, offsets, counts):
        ffi, lib = self.ffi, self.lib
        dp = ffi.new("void*[]", [ffi.cast("void*", p) for p in dst_ptrs])
        off = ffi.new("unsigned long long[]", [int(o) for o in offsets])
        cnt = ffi.new("unsigned long long[]", [int(c) for c in counts])
        lib.gdfx_xjoin_scatter(self.C.Column(keys).cdata, id_base, ranks, nlocal, dp, off, cnt)

    def xjoin_count_dev(self, keys, ranks, nlocal, d_counts):
        self.lib.gdfx_xjoin_count_dev(self.C.Column(keys).cdata, ranks, nlocal, self.ffi.cast("unsigned long long*", d_counts.data_ptr()))

    def xjoin_plan_dev(self, d_all, ranks, nlocal, rank, cap_build, cap_probe, d_off_build, d_off_probe, d_status):
        c = self.ffi.cast
        self.lib.gdfx_xjoin_plan_dev(c("unsigned long long*", d_all.data_ptr()), ranks, nlocal, rank, cap_build, cap_probe,
                                     c("unsigned long long*", d_off_build.data_ptr()), c("unsigned long long*", d_off_probe.data_ptr()),
                                     c("int*", d_status.data_ptr()))

    def xjoin_scatter_dev(self, keys, id_base, ranks, nlocal, dst_ptrs, d_offsets, d_counts, d_status, ctas_per_sm=0):
        ffi, lib = self.ffi, self.lib
        dp = ffi.new("void*[]", [ffi.cast("void*", p) for p in dst_ptrs])
        lib.gdfx_xjoin_scatter_dev(self.C.Column(keys).cdata, id_base, ranks, nlocal, dp,
                                   ffi.cast("unsigned long long*", d_offsets.data_ptr()),
                                   ffi.cast("unsigned long long*", d_counts.data_ptr()), ffi.cast("int*", d_status.data_ptr()),
                                   ctas_per_sm)

    def xjoin_build(self, build_ptr, build_counts, nlocal, overlap):
        ffi, lib = self.ffi, self.lib
        h = ffi.new("void**")
        lib.gdfx_xjoin_build(ffi.cast("void*", build_ptr), ffi.new("unsigned long long[]", [int(c) for c in build_counts]), nlocal,
                             1 if overlap else 0, h)
        return h[0]

    def xjoin_probe(self, handle, probe_ptr, probe_counts):
        C, ffi, lib = self.C, self.ffi, self.lib
        out_l, out_r = ffi.new("gdf_column*"), ffi.new("gdf_column*")
        lib.gdfx_xjoin_probe(handle, ffi.cast("void*", probe_ptr), ffi.new("unsigned long long[]", [int(c) for c in probe_counts]),
                             out_l, out_r)
        return C.library_owned_view(out_l), C.library_owned_view(out_r)

    def xjoin_local(self, probe_ptr, probe_counts, build_ptr, build_counts, nlocal):
        C, ffi, lib = self.C, self.ffi, self.lib
        out_l, out_r = ffi.new("gdf_column*"), ffi.new("gdf_column*")
        lib.gdfx_xjoin_local(ffi.cast("void*", probe_ptr), ffi.new("unsigned long long[]", [int(c) for c in probe_counts]),
                             ffi.cast("void*", build_ptr), ffi.new("unsigned long long[]", [int(c) for c in build_counts]),
                             nlocal, out_l, out_r)
        return C.library_owned_view(out_l), C.library_owned_view(out_r)

    def peer_alloc(self, nbytes):
        ffi, lib = self.ffi, self.lib
        ptr = ffi.new("void**")
        handle = ffi.new("char[64]")
        lib.gdfx_peer_alloc(ptr, nbytes, handle)
        return int(ffi.cast("uintptr_t", ptr[0])), bytes(ffi.buffer(handle, 64))

    def peer_open(self, handle):
        ffi, lib = self.ffi, self.lib
        ptr = ffi.new("void**")
        lib.gdfx_peer_open(ffi.new("char[64]", handle), ptr)
        return int(ffi.cast("uintptr_t", ptr[0]))

    def peer_close(self, address):
        self.lib.gdfx_peer_close(self.ffi.cast("void*", address))

    def peer_free(self, address):
        self.lib.gdfx_peer_free(self.ffi.cast("void*", address))

    def view(self, address, nelem, np_dtype):
        return self.C._alias(address, nelem, np_dtype) if nelem else torch.empty(0, dtype=getattr(torch, np.dtype(np_dtype).name), device="cuda")

    def _join(self, kind, lkeys, rkeys, lids, rids):
        """Join of exchanged {key, id} pairs; the outputs are the ids (the pairs' row tags ARE the ids
        inside the partitioned join, so no separate index->id gather pass is needed)."""
        C, ffi, lib = self.C, self.ffi, self.lib
        L, R = C.Column(lkeys), C.Column(rkeys)
        out_l, out_r = ffi.new("gdf_column*"), ffi.new("gdf_column*")
        lib.gdfx_join_pairs(kind, L.cdata, ffi.cast("int32_t*", lids.data_ptr()), R.cdata,
                            ffi.cast("int32_t*", rids.data_ptr()), out_l, out_r)
        return C.library_owned_to_torch(out_l), C.library_owned_to_torch(out_r)

    def inner_join(self, lkeys, rkeys, lids, rids):
        return self._join(0, lkeys, rkeys, lids, rids)

    def left_join(self, lkeys, rkeys, lids, rids):
        return self._join(1, lkeys, rkeys, lids, rids)

    def group_by_sum(self, keys, vals):
        C, ffi, lib = self.C, self.ffi, self.lib
        n = keys.numel()
        if n == 0:
            return keys.clone(), vals.clone()
        K, V = C.Column(keys), C.Column(vals)
        OK, OV = C.Column(torch.empty_like(keys)), C.Column(torch.empty_like(vals))
        lib.gdf_group_by_sum(1, C.column_array([K]), V.cdata, ffi.NULL, C.column_array([OK]), OV.cdata, self.ctx)
        g = int(OV.cdata.size)
        return OK.data[:g], OV.data[:g]


# ------------------------------------------------------------------------------------------------
# the exchange step (backend-agnostic: nccl on GPUs, gloo in the CPU unit tests)
# ------------------------------------------------------------------------------------------------
def exchange(cols, offsets, group=None):
    """Personalised all-to-all of partition-contiguous columns.

    cols     list of 1-D tensors, all with n rows, partition p occupying rows [offsets[p], offsets[p+1])
    offsets  world_size start offsets (the host ``partition_offsets`` array of gdf_hash_partition)
    Returns (received columns, recv_counts): the rows every rank sent to this one, source-rank major.
    One int64 counts exchange + ONE all_to_all_single per column.
    """
    world = dist.get_world_size(group)
    n = cols[0].numel()
    bounds = list(offsets) + [n]
    send_counts = [int(bounds[p + 1] - bounds[p]) for p in range(world)]
    dev = cols[0].device
    sc = torch.tensor(send_counts, dtype=torch.int64, device=dev)
    rc = torch.empty(world, dtype=torch.int64, device=dev)
    dist.all_to_all_single(rc, sc, group=group)
    recv_counts = [int(x) for x in rc.tolist()]
    total = sum(recv_counts)
    outs = []
    for c in cols:
        out = torch.empty(total, dtype=c.dtype, device=dev)
        dist.all_to_all_single(out, c.contiguous(), recv_counts, send_counts, group=group)
        outs.append(out)
    return outs, recv_counts


class PeerExchange(object):
    """Fused partition + exchange: the partition kernel of every rank writes its {key, id} pairs straight
    into the destination ranks' receive buffers through NVLink peer mappings (CUDA IPC), so the rows cross
    the fabric from inside the scatter kernel and no separate all-to-all pass (read + send + write) exists.

    Per exchange: one histogram pass, one all_gather of the world x world count matrix (which is also the
    "everybody is done reading the previous contents" barrier, being stream-ordered after each rank's
    previous work), one scatter pass, one 4-byte all_reduce ("all stores have landed").
    Receive buffers are cudaMalloc'd once per name and grown only when a call needs more rows.
    """

    def __init__(self, ops, group=None):
        self.ops, self.group = ops, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.slots = {}   # name -> {"cap", "itemsize", "mine": (kptr, iptr), "peers": ([kptr..], [iptr..])}
        self._flag = None
        self.one_pass = True                  # fused_inner_join (one partition pass per side) before the two-pass path
        self.async_plan = True                # steady state keeps counts / plan on the device (no host round trips)
        self._async_state = {}
        self.use_symm_mem = False             # receive buffers from torch symmetric memory instead of cudaMalloc + CUDA IPC
        self.overlap_build = True             # fill the hash tables on a second stream during the probe side's exchange
        self.scatter_ctas_per_sm = 2          # ... whose scatter then leaves a third of every SM to the table build
        self.rows_per_partition = 1 << 20     # build rows per receiver-local partition of the one-pass exchange

    @staticmethod
    def available(ops, group=None):
        """Collective self-test: can every rank map every other rank's memory (CUDA IPC + peer access)?
        All ranks get the same answer, so callers can fall back to the NCCL exchange together."""
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        ok, opened, ptr = 1, [], None
        try:
            ptr, handle = ops.peer_alloc(256)
        except Exception:
            ok, handle = 0, b""
        handles = [None] * world
        dist.all_gather_object(handles, handle, group=group)
        if ok:
            try:
                for r, h in enumerate(handles):
                    if r != rank:
                        if len(h) != 64:
                            raise RuntimeError("peer %d has no handle" % r)
                        opened.append(ops.peer_open(h))
            except Exception:
                ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        for p in opened:
            try:
                ops.peer_close(p)
            except Exception:
                pass
        dist.barrier(group=group)
        if ptr is not None:
            ops.peer_free(ptr)
        return bool(flag.item())

    def _ensure(self, name, rows, itemsize):
        slot = self.slots.get(name)
        if slot and slot["cap"] >= rows and slot["itemsize"] == itemsize:
            return slot
        if slot:
            self._release(slot)
        cap = max(1 << 20, (rows + (rows >> 3) + (1 << 20) - 1) >> 20 << 20)   # 12.5 % slack, 1 Mi-row granules
        kptr, kh = self.ops.peer_alloc(cap * itemsize)
        iptr, ih = self.ops.peer_alloc(cap * 4)
        handles = [None] * self.world
        dist.all_gather_object(handles, (kh, ih), group=self.group)
        pk, pi = [], []
        for r, (hk, hi) in enumerate(handles):
            if r == self.rank:
                pk.append(kptr), pi.append(iptr)
            else:
                pk.append(self.ops.peer_open(hk)), pi.append(self.ops.peer_open(hi))
        slot = {"cap": cap, "itemsize": itemsize, "mine": (kptr, iptr), "peers": (pk, pi)}
        self.slots[name] = slot
        return slot

    def _release(self, slot):
        torch.cuda.synchronize()
        dist.barrier(group=self.group)
        if slot.get("symm"):
            slot.pop("symm")
            return
        for r in range(self.world):
            if r != self.rank:
                for peers in slot["peers"]:
                    self.ops.peer_close(peers[r])
        dist.barrier(group=self.group)
        for ptr in slot["mine"]:
            self.ops.peer_free(ptr)

    def close(self):
        for slot in self.slots.values():
            self._release(slot)
        self.slots = {}

    def _ensure_pairs(self, name, rows):
        """One IPC-shared buffer of 8-byte {key32, id32} pairs per name (fused one-pass exchange)."""
        slot = self.slots.get(name)
        if slot and slot["cap"] >= rows and slot.get("pairs"):
            return slot
        if slot:
            self._release(slot)
        cap = max(1 << 20, (rows + (rows >> 3) + (1 << 20) - 1) >> 20 << 20)
        if self.use_symm_mem:    # torch symmetric memory (CUDA VMM allocations exchanged as file descriptors) instead of legacy IPC
            import torch.distributed._symmetric_memory as symm
            t = symm.empty(cap * 2, dtype=torch.int32, device=torch.device("cuda", torch.cuda.current_device()))
            hdl = symm.rendezvous(t, group=self.group if self.group is not None else dist.group.WORLD)
            peers = [int(p) for p in hdl.buffer_ptrs]
            slot = {"cap": cap, "itemsize": 8, "pairs": True, "mine": (int(t.data_ptr()),), "peers": (peers,), "symm": (t, hdl)}
            self.slots[name] = slot
            return slot
        ptr, h = self.ops.peer_alloc(cap * 8)
        handles = [None] * self.world
        dist.all_gather_object(handles, h, group=self.group)
        peers = [ptr if r == self.rank else self.ops.peer_open(hh) for r, hh in enumerate(handles)]
        slot = {"cap": cap, "itemsize": 8, "pairs": True, "mine": (ptr,), "peers": (peers,)}
        self.slots[name] = slot
        return slot

    def fused_inner_join(self, probe_keys, build_keys, probe_offset, build_offset, timings=None, global_build_rows=None):
        """INNER join with ONE partition pass per side: every rank histograms both sides with the combined
        (destination rank x receiver-local partition) geometry, ONE all_gather carries both count matrices, ONE
        scatter per side stores compact pairs into the peers' partition-contiguous buffers over NVLink, and the
        receiver joins what it got without partitioning again.  Returns None when the compact form does not apply
        (a build key wider than 32 bits, a partition too large): the caller then takes the two-pass path.

        Steady state (receive buffers exist) is ASYNCHRONOUS: counts, the all-gathered matrix and the write offsets stay
        on the device (gdfx_xjoin_count_dev / _plan_dev / _scatter_dev), a side stream copies the matrix to pinned host
        memory while the scatters run, and the host only reads it when it needs the per-partition totals for the local
        join.  The device-side plan raises a flag instead of writing when a buffer would overflow or a build key is wide;
        the host sees the flag afterwards and falls back to the synchronous route, which also (re)allocates buffers."""
        world, rank, dev = self.world, self.rank, probe_keys.device
        ev = _Stamps(timings, dev)
        if global_build_rows is None:       # one collective + host read; callers that know their table sizes pass them
            sizes = torch.tensor([build_keys.numel()], dtype=torch.int64, device=dev)
            dist.all_reduce(sizes, group=self.group)
            global_build_rows = int(sizes.item())
        nlocal = fused_nlocal(global_build_rows, world, self.rows_per_partition)
        bins = world * nlocal
        sb, sp = self.slots.get("xbuild"), self.slots.get("xprobe")
        if self.async_plan and sb and sp and sb.get("pairs") and sp.get("pairs"):
            out = self._fused_async(probe_keys, build_keys, probe_offset, build_offset, nlocal, sb, sp, ev)
            if out is not False:
                return out
        cb, hi_b = self.ops.xjoin_count(build_keys, world, nlocal)
        cp, _ = self.ops.xjoin_count(probe_keys, world, nlocal)
        mine = torch.tensor(cb + cp + [hi_b], dtype=torch.int64, device=dev)
        allc = torch.empty(world * (2 * bins + 1), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(allc, mine, group=self.group)     # also: everybody is done with the previous buffers
        M = allc.view(world, 2 * bins + 1).cpu().tolist()
        if any(row[2 * bins] for row in M):
            return None
        off_b, tot_b, recv_b = plan_fused_exchange([row[:bins] for row in M], world, nlocal, rank)
        off_p, tot_p, recv_p = plan_fused_exchange([row[bins:2 * bins] for row in M], world, nlocal, rank)
        if max(tot_b) > (1 << 22):
            return None
        sb, sp = self._ensure_pairs("xbuild", max(recv_b)), self._ensure_pairs("xprobe", max(recv_p))
        ev.mark("count+plan")
        self.ops.xjoin_scatter(build_keys, build_offset, world, nlocal, sb["peers"][0], off_b, cb)
        self.ops.xjoin_scatter(probe_keys, probe_offset, world, nlocal, sp["peers"][0], off_p, cp)
        if self._flag is None:
            self._flag = torch.zeros(1, dtype=torch.int32, device=dev)
        dist.all_reduce(self._flag, group=self.group)                 # stream-ordered: every rank's stores have landed
        ev.mark("partition+exchange")
        out = self.ops.xjoin_local(sp["mine"][0], tot_p, sb["mine"][0], tot_b, nlocal)
        ev.mark("local_join")
        return out

    def _fused_async(self, probe_keys, build_keys, probe_offset, build_offset, nlocal, sb, sp, ev):
        """The steady-state route of fused_inner_join.  Returns False when the synchronous route has to run instead
        (a receive buffer is too small), None / (left, right) otherwise."""
        world, rank, dev = self.world, self.rank, probe_keys.device
        bins = world * nlocal
        stride = 2 * (bins + 1)                      # per rank: build bins | build hi_or | probe bins | probe hi_or
        st = self._async_state.get(stride)
        if st is None:
            st = {"mine": torch.zeros(stride, dtype=torch.int64, device=dev),
                  "all": torch.zeros(world * stride, dtype=torch.int64, device=dev),
                  "off": torch.zeros(2 * bins, dtype=torch.int64, device=dev),
                  "status": torch.zeros(2, dtype=torch.int32, device=dev),
                  "pin": torch.zeros(world * stride + 2, dtype=torch.int64).pin_memory(),
                  "side": torch.cuda.Stream(device=dev)}
            self._async_state[stride] = st
        mine, allc, off, status = st["mine"], st["all"], st["off"], st["status"]
        self.ops.xjoin_count_dev(build_keys, world, nlocal, mine[:bins + 1])
        self.ops.xjoin_count_dev(probe_keys, world, nlocal, mine[bins + 1:])
        dist.all_gather_into_tensor(allc, mine, group=self.group)     # also: everybody is done with the previous buffers
        self.ops.xjoin_plan_dev(allc, world, nlocal, rank, sb["cap"], sp["cap"], off[:bins], off[bins:], status)
        gathered = torch.cuda.Event()
        gathered.record()
        with torch.cuda.stream(st["side"]):         # matrix + flags to pinned memory while the scatters run
            st["side"].wait_event(gathered)
            st["pin"][:world * stride].copy_(allc, non_blocking=True)
            st["pin"][world * stride:].copy_(status.to(torch.int64), non_blocking=True)
            copied = torch.cuda.Event()
            copied.record()
        ev.mark("count+plan")
        if self._flag is None:
            self._flag = torch.zeros(1, dtype=torch.int32, device=dev)
        self.ops.xjoin_scatter_dev(build_keys, build_offset, world, nlocal, sb["peers"][0], off[:bins], mine[:bins], status)
        dist.all_reduce(self._flag, group=self.group)                 # stream-ordered: every rank's BUILD pairs have landed
        copied.synchronize()                                          # long done: the copy only waited for the all_gather
        M = st["pin"][:world * stride].view(world, stride).numpy()
        wide, overflow = int(st["pin"][world * stride]), int(st["pin"][world * stride + 1])
        if wide or overflow:                                          # nothing was / will be written (the scatters check the flags)
            return None if wide else False                            # two-pass path / synchronous route that grows the buffers
        tot_b = _granules(M[:, rank * nlocal:(rank + 1) * nlocal]).sum(0)          # slots are padded (plan_fused_exchange)
        tot_p = _granules(M[:, bins + 1 + rank * nlocal:bins + 1 + (rank + 1) * nlocal]).sum(0)
        if int(tot_b.max()) > (1 << 22):
            return None
        # tables are filled on a private stream while the probe side crosses NVLink on this one
        handle = self.ops.xjoin_build(sb["mine"][0], tot_b.tolist(), nlocal, self.overlap_build)
        self.ops.xjoin_scatter_dev(probe_keys, probe_offset, world, nlocal, sp["peers"][0], off[bins:], mine[bins + 1:2 * bins + 1],
                                   status, self.scatter_ctas_per_sm if self.overlap_build else 0)
        dist.all_reduce(self._flag, group=self.group)                 # every rank's PROBE pairs have landed
        ev.mark("partition+exchange")
        out = self.ops.xjoin_probe(handle, sp["mine"][0], tot_p.tolist())
        ev.mark("local_join")
        return out

    def exchange_pairs(self, name, keys, id_base):
        """Returns (keys, ids) this rank received: zero-copy views of its receive buffers, valid until the
        next exchange under the same name."""
        world, rank, dev = self.world, self.rank, keys.device
        mine = torch.tensor(self.ops.partition_count(keys, world), dtype=torch.int64, device=dev)
        allc = torch.empty(world * world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(allc, mine, group=self.group)
        C = allc.view(world, world).cpu()                    # C[s][d] = rows rank s sends to rank d
        recv_totals = C.sum(0)
        slot = self._ensure(name, int(recv_totals.max().item()), keys.element_size())
        offsets = [int(C[:rank, d].sum().item()) for d in range(world)]
        self.ops.partition_scatter_peer(keys, id_base, slot["peers"][0], slot["peers"][1], offsets)
        if self._flag is None:
            self._flag = torch.zeros(1, dtype=torch.int32, device=dev)
        dist.all_reduce(self._flag, group=self.group)        # stream-ordered: every rank's stores have landed
        n = int(recv_totals[rank].item())
        np_key = {8: np.int64, 4: np.int32}[keys.element_size()]
        return self.ops.view(slot["mine"][0], n, np_key), self.ops.view(slot["mine"][1], n, np.int32)


EXCHANGE_GRANULE = 4      # pairs per 32-byte sector: csrc/join_part.cu kXG


def _granules(c):
    """Counts rounded up to whole exchange granules (numpy array or int)."""
    return (c + EXCHANGE_GRANULE - 1) & ~(EXCHANGE_GRANULE - 1)


def plan_fused_exchange(counts, world, nlocal, rank):
    """Host-side plan of the one-pass exchange.  counts[s][d * nlocal + p] = pairs rank s sends to (rank d, local
    partition p).  Receiver d lays its buffer out partition-major, and inside a partition in sender order, so
        offsets[d * nlocal + p]  where THIS rank's pairs of bin (d, p) start inside d's buffer
        part_totals[p]           pairs of local partition p this rank receives (all senders, pads included)
        recv_rows[d]             total pairs rank d receives (buffer sizing)
    Slots are padded to whole 32-byte granules so that the scatter sends nothing but sector-aligned bulk stores over NVLink.
    Pure arithmetic on the gathered count matrix: every rank computes the same plan (tests/test_dist_cpu.py)."""
    c = np.asarray(counts, dtype=np.int64).reshape(world, world * nlocal)
    c = _granules(c)      # every (sender, bin) slot holds whole 32-byte granules: the rest is padded with "no row" pairs
    totals = c.sum(0)                                             # pairs of every bin, all senders
    before_me = c[:rank].sum(0)                                   # ... of the senders ranked before this one
    bin_start = (np.cumsum(totals.reshape(world, nlocal), 1) - totals.reshape(world, nlocal)).reshape(-1)   # inside its destination
    offsets = (bin_start + before_me).tolist()
    recv_rows = totals.reshape(world, nlocal).sum(1).tolist()
    part_totals = totals[rank * nlocal:(rank + 1) * nlocal].tolist()
    return offsets, part_totals, recv_rows


def fused_nlocal(build_rows_total, world, rows_per_partition=1 << 20, max_bins=256):
    """Receiver-local radix partitions for the one-pass exchange: a power of two such that one partition holds about
    2^20 build rows (table <= 16 MB, L2-resident), limited by the 256 bins one partition pass can write."""
    per_rank = (build_rows_total + world - 1) // world
    want = max(1, (per_rank + rows_per_partition - 1) // rows_per_partition)
    n = 1
    while n < want:
        n *= 2
    while n * world > max_bins and n > 1:
        n //= 2
    return n


def shard_bounds(total_rows, world, rank):
    """Block distribution of total_rows over world ranks: [lo, hi) of `rank`."""
    per = (total_rows + world - 1) // world
    lo = min(total_rows, rank * per)
    return lo, min(total_rows, lo + per)


# ------------------------------------------------------------------------------------------------
# distributed operators
# ------------------------------------------------------------------------------------------------
def distributed_join(kind, left_keys, right_keys, left_offset, right_offset, ops, group=None, timings=None, peer=None,
                     global_rows=None):
    """Hash join of block-distributed key columns.

    left_keys / right_keys   this rank's shard of the (single, integer) key column
    left_offset/right_offset global row index of the shard's first row
    global_rows              optional (total left rows, total right rows) over all ranks: a caller that knows its table
                             sizes saves the collective + host read that would otherwise establish them
    Returns (left_idx, right_idx): int32 GLOBAL row ids of this rank's share of the result (-1 =
    no partner, LEFT join).  The union over ranks is the join of the full tables (pair order
    unspecified, as in the reference).
    """
    world = dist.get_world_size(group)
    dev = left_keys.device
    ev = _Stamps(timings, dev)
    if max(left_offset + left_keys.numel(), right_offset + right_keys.numel()) >= 2 ** 31:
        raise ValueError("global row ids must fit int32 (gdf join indices are GDF_INT32)")
    fn = ops.inner_join if kind == "inner" else ops.left_join
    if peer is not None and kind == "inner" and getattr(peer, "one_pass", True):
        # one partition pass per side: the build side is the smaller table (gdf_inner_join's rule, joining.h:59-67)
        if global_rows is None:
            sizes = torch.tensor([left_keys.numel(), right_keys.numel()], dtype=torch.int64, device=dev)
            dist.all_reduce(sizes, group=group)
            global_rows = tuple(int(x) for x in sizes.tolist())
        flip = global_rows[1] > global_rows[0]
        out = (peer.fused_inner_join(right_keys, left_keys, right_offset, left_offset, timings, global_rows[0]) if flip
               else peer.fused_inner_join(left_keys, right_keys, left_offset, right_offset, timings, global_rows[1]))
        if out is not None:
            return (out[1], out[0]) if flip else out
    if peer is not None:   # two-pass path: partition + exchange over NVLink peer memory, then the local partitioned join
        lk, li = peer.exchange_pairs("left", left_keys, left_offset)
        rk, ri = peer.exchange_pairs("right", right_keys, right_offset)
        ev.mark("partition+exchange")
        out = fn(lk, rk, li, ri)
        ev.mark("local_join")
        return out
    lk, li, loff = ops.partition_pairs(left_keys, left_offset, world)
    rk, ri, roff = ops.partition_pairs(right_keys, right_offset, world)
    ev.mark("partition")
    (lk, li), _ = exchange([lk, li], loff, group)
    (rk, ri), _ = exchange([rk, ri], roff, group)
    ev.mark("all_to_all")
    out = fn(lk, rk, li, ri)
    ev.mark("local_join")
    return out


def distributed_left_join_masked(left_keys, left_valids, right_keys, right_valids, left_offset, right_offset, ops,
                                 group=None, timings=None, peer=None):
    """LEFT hash join on a COMPOSITE key with NULLs (BASELINE config C5: (int64,int32) key, 30 % null rows).

    left_keys / right_keys     lists of this rank's key-column shards
    left_valids / right_valids matching lists of packed validity masks (torch.uint8, LSB first) or None
    Row validity (AND over the key columns) is turned into a byte column that travels with the rows:
    {key columns, global row id, row-valid byte} are hash-partitioned on the key columns
    (``gdf_hash_partition``), exchanged with one all_to_all per column, and joined per rank with
    ``gdf_left_join`` on the rebuilt mask - so the reference's null rule holds unchanged across ranks:
    a left row with a NULL key yields (l, -1), a right row with a NULL key never matches
    (ref join_kernels.cuh:59,161,316).
    """
    world = dist.get_world_size(group)
    dev = left_keys[0].device
    ev = _Stamps(timings, dev)
    nk = len(left_keys)
    lids = torch.arange(left_offset, left_offset + left_keys[0].numel(), dtype=torch.int32, device=dev)
    rids = torch.arange(right_offset, right_offset + right_keys[0].numel(), dtype=torch.int32, device=dev)
    lok = ops.rows_valid_bytes(left_keys, left_valids)
    rok = ops.rows_valid_bytes(right_keys, right_valids)
    lcols, rcols = list(left_keys) + [lids, lok], list(right_keys) + [rids, rok]
    if world > 1:
        lcols, loff = ops.hash_partition_rows(lcols, nk, world)
        rcols, roff = ops.hash_partition_rows(rcols, nk, world)
        ev.mark("partition")
        lcols, _ = exchange(lcols, loff, group)
        rcols, _ = exchange(rcols, roff, group)
        ev.mark("all_to_all")
    out = ops.left_join_masked(lcols[:nk], lcols[nk + 1], rcols[:nk], rcols[nk + 1], lcols[nk], rcols[nk])
    ev.mark("local_join")
    return out


def distributed_group_by_sum(keys, vals, ops, group=None, timings=None):
    """Two-phase hash group-by SUM of block-distributed (key, value) rows.  Returns this rank's share
    of the groups: every distinct key appears on exactly one rank."""
    world = dist.get_world_size(group)
    ev = _Stamps(timings, keys.device)
    pk, pv = ops.group_by_sum(keys, vals)          # phase 1: local partials (<= #groups rows)
    ev.mark("local_groupby")
    if world == 1:
        return pk, pv
    (pk, pv), off = ops.hash_partition([pk, pv], world)
    ev.mark("partition")
    (pk, pv), _ = exchange([pk, pv], off, group)
    ev.mark("all_to_all")
    out = ops.group_by_sum(pk, pv)                 # phase 2: merge the partials this rank owns
    ev.mark("merge")
    return out


class _Stamps(object):
    """Optional per-phase device timing (CUDA events on the current stream)."""

    def __init__(self, sink, device):
        self.sink = sink
        self.cuda = sink is not None and device.type == "cuda"
        self.last = self._now() if self.cuda else None

    def _now(self):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    def mark(self, name):
        if not self.cuda:
            return
        e = self._now()
        self.sink.setdefault("_events", []).append((name, self.last, e))
        self.last = e


def resolve_timings(sink):
    """After a synchronize: fold recorded event pairs into {phase: ms}."""
    out = {}
    for name, a, b in sink.pop("_events", []):
        out[name] = out.get(name, 0.0) + a.elapsed_time(b)
    return out
