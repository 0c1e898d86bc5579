"""Multi-GPU BFS and SSSP: 1-D vertex partition, one process per GPU, per-level exchange.

The reference has no multi-GPU path (its operators throw when ``context.size() != 1``,
include/gunrock/framework/operators/advance/advance.hxx:125-128).  This module is the
extension BASELINE.json asks for: rank r owns the contiguous vertex range [r*n/P, (r+1)*n/P) — its CSR rows
with GLOBAL column ids and its slice of the result array.

BFS: the frontier / visited sets are replicated 1-bit-per-vertex maps (32 MiB at scale-28), so an exchange is a
fixed-size collective:

  top-down level   local advance ORs unvisited neighbours into a global-length candidate map
                   -> all_to_all of the P candidate slices, owner ORs them        (n/8 bytes sent per rank)
  bottom-up level  owner walks the in-edges of its unvisited vertices against the replicated frontier map
                   (no candidate exchange at all)
  both             owner absorbs candidates (depth, visited), then ONE all_gather carries the new frontier
                   slice together with the two Beamer counters, so direction choice and termination are
                   decided identically on every rank with no extra all_reduce.

SSSP: a replicated array of tentative distances; per round relax -> reduce_scatter(min) to the owners ->
owner-side collect of the rows that improved (PartitionedSSSP below).

Three drivers of the same logic:
  * NativePartitionedBFS (.bfs / .sssp) — the product: whole loops in C++ (ess_dist_bfs / ess_dist_sssp,
    essentials_b200/csrc/capi_dist.cu); the exchange runs in the library's own peer-memory kernels over NVLink
    (IPC-mapped windows + epoch flags), NCCL bootstraps and is the fallback.
  * PartitionedBFS / PartitionedSSSP with CudaBackend — the same loops over torch.distributed, calling the per-level
    kernels through the C ABI (ess_bfs_partition_step / _pull / ess_bfs_absorb / ess_sssp_partition_*).
  * the same classes with a numpy backend on the gloo backend — tests/test_dist_gloo.py, CPU only; the stand-in lives
    in tests/, there is no CPU fallback in the product.
"""
from __future__ import annotations

import ctypes
from ctypes import byref, c_int64

import torch
import torch.distributed as dist

INF = 2**31 - 1


class CudaBackend:
    """Local level kernels of one rank, in libessentials_b200.so. Both calls only enqueue work."""

    def __init__(self, csr_local, row_begin: int, n_global: int, device, stream):
        import essentials_b200 as ess
        self.ess = ess
        self.ctx = ess.Context(device.index, stream=stream)
        self.graph = ess.Graph(csr_local, symmetric=True, partition=True)
        self.row_begin, self.n_global = row_begin, n_global
        # bottom-up hints need the degree of remote neighbours: gather all degrees once at setup
        deg_local = (csr_local.offsets[1:] - csr_local.offsets[:-1]).to(torch.int32).contiguous()
        deg_global = torch.empty(n_global, dtype=torch.int32, device=device)
        dist.all_gather_into_tensor(deg_global, deg_local)
        self.graph.build_pull_hints(deg_global)
        del deg_global

    def step(self, pull: bool, frontier_bits, visited_bits, candidate_bits, frontier_list, frontier_count: int):
        ess = self.ess
        ess._check(ess.lib().ess_bfs_partition_step(self.ctx.handle, self.graph.handle, self.row_begin, self.n_global,
                                                    int(pull), ess._p(frontier_bits), ess._p(visited_bits),
                                                    ess._p(candidate_bits), ess._p(frontier_list),
                                                    int(frontier_count)), "ess_bfs_partition_step")

    def pull(self, level: int, frontier_bits, visited_bits, next_slice, depth_local, counts):
        ess = self.ess
        ess._check(ess.lib().ess_bfs_partition_pull(self.ctx.handle, self.graph.handle, self.row_begin, level,
                                                    ess._p(frontier_bits), ess._p(visited_bits), ess._p(next_slice),
                                                    ess._p(depth_local), ess._p(counts)), "ess_bfs_partition_pull")

    def gather_fresh(self, next_slice, fresh_list):
        """Sparse list (local row ids) of the bits of the owned next-frontier slice; enqueue only."""
        ess = self.ess
        ess._check(ess.lib().ess_bits_to_list_async(self.ctx.handle, ess._p(next_slice), int(next_slice.numel()) * 32,
                                                    ess._p(fresh_list)), "ess_bits_to_list_async")

    def merge(self, gathered, world: int, slice_words: int, frontier_bits, visited_bits, counts_out):
        ess = self.ess
        ess._check(ess.lib().ess_bfs_merge_gathered(self.ctx.handle, ess._p(gathered), world, slice_words,
                                                    ess._p(frontier_bits), ess._p(visited_bits), ess._p(counts_out)),
                   "ess_bfs_merge_gathered")

    def absorb(self, level: int, candidates, n_slices: int, stride_words: int, visited_bits, next_slice, depth_local,
               fresh_list, counts):
        ess = self.ess
        ess._check(ess.lib().ess_bfs_absorb(self.ctx.handle, self.graph.handle, self.row_begin, level,
                                            ess._p(candidates), n_slices, stride_words, ess._p(visited_bits),
                                            ess._p(next_slice), ess._p(depth_local), ess._p(fresh_list),
                                            ess._p(counts)), "ess_bfs_absorb")

    # ---- SSSP ----
    def relax(self, active_list, active_count: int, dist_local, replica):
        ess = self.ess
        ess._check(ess.lib().ess_sssp_partition_relax(self.ctx.handle, self.graph.handle, ess._p(active_list),
                                                      int(active_count), ess._p(dist_local), ess._p(replica)),
                   "ess_sssp_partition_relax")

    def collect(self, reduced, dist_local, active_list, counts):
        ess = self.ess
        ess._check(ess.lib().ess_sssp_partition_collect(self.ctx.handle, self.graph.handle, ess._p(reduced),
                                                        ess._p(dist_local), ess._p(active_list), ess._p(counts)),
                   "ess_sssp_partition_collect")

    def launches(self) -> int:
        return self.ctx.launches()


class PartitionedBFS:
    """Direction-optimising BFS over a 1-D partitioned symmetric graph."""

    def __init__(self, csr_local, row_begin: int, n_global: int, rank: int, world: int, backend, device,
                 alpha: float = 14.0, beta: float = 24.0):
        assert n_global % (64 * world) == 0, "partition boundaries must be multiples of 64 vertices"
        self.rank, self.world, self.device, self.backend = rank, world, device, backend
        self.n_global, self.per = n_global, n_global // world
        self.row_begin = row_begin
        assert row_begin == rank * self.per and csr_local.offsets.numel() - 1 == self.per
        self.csr = csr_local
        self.alpha, self.beta = alpha, beta
        self.words = n_global // 32
        self.wper = self.per // 32
        i32 = dict(dtype=torch.int32, device=device)
        self.frontier_bits = torch.zeros(self.words, **i32)
        self.visited_bits = torch.zeros(self.words + 1, **i32)
        self.candidate_bits = torch.zeros(self.words + 1, **i32)
        self.depth_local = torch.empty(self.per, **i32)
        self.fresh_list = torch.zeros(self.per, **i32)  # sparse frontier of this rank (local row ids)
        # payload of the per-level all_gather: the owned next-frontier slice + (|fresh|, Σdeg fresh) as 2 x int64,
        # both written by the absorb kernel
        self.send = torch.zeros(self.wper + 4, **i32)
        self.recv = torch.zeros(world * (self.wper + 4), **i32)
        self.a2a_recv = torch.zeros(self.words, **i32)
        self.counts_dev = torch.zeros(2 * world, dtype=torch.int64, device=device)
        self.counts_host = torch.zeros(2 * world, dtype=torch.int64)
        if device.type == "cuda":
            self.counts_host = self.counts_host.pin_memory()
        self.deg_local = (csr_local.offsets[1:] - csr_local.offsets[:-1]).to(torch.int64)
        m = torch.tensor([int(csr_local.indices.numel())], dtype=torch.int64, device=device)
        dist.all_reduce(m)
        self.m_global = int(m.item())
        self.offset_bits = 64 if csr_local.offsets.element_size() == 8 else 32
        # isolated vertices never join a frontier: mark them visited once (replicated map)
        iso_local = _pack_bits(self.deg_local == 0)
        iso = torch.empty(self.words, **i32)
        dist.all_gather_into_tensor(iso, iso_local)
        self.isolated_bits = iso
        self.levels = self.pull_levels = 0
        self.bytes_exchanged = 0

    # ------------------------------------------------------------------------------------------------
    def _sync(self):
        if self.device.type == "cuda":
            torch.cuda.current_stream(self.device).synchronize()

    def owner(self, v: int) -> int:
        return v // self.per

    def bfs(self, source: int, trace: list | None = None) -> dict:
        """Runs one BFS; depths of the owned range end up in self.depth_local. Returns per-run statistics.
        One host synchronisation per level: reading the counters that arrive with the all_gather.
        `trace` (development aid): a list that receives (level, phase, host seconds) tuples, with a device
        synchronisation after every phase."""
        import time as _time

        def mark(phase, t0):
            if trace is None:
                return t0
            torch.cuda.synchronize() if self.device.type == "cuda" else None
            t1 = _time.perf_counter()
            trace.append((level, phase, t1 - t0))
            return t1
        dev = self.device
        self.visited_bits[: self.words].copy_(self.isolated_bits)
        self.frontier_bits.zero_()
        self.depth_local.fill_(INF)
        word, bit = source // 32, source % 32
        mask = _bit(bit)
        self.frontier_bits[word] = mask
        self.visited_bits[word] |= mask
        local = torch.zeros(2, dtype=torch.int64, device=dev)
        my_count = 0
        if self.owner(source) == self.rank:
            self.depth_local[source - self.row_begin] = 0
            self.fresh_list[0] = source - self.row_begin
            my_count = 1
            local[0] = 1
            local[1] = self.deg_local[source - self.row_begin]
        dist.all_reduce(local)
        n_f, m_f = (int(x) for x in local.tolist())
        m_u = self.m_global - m_f
        prev_n_f, pulling, level, pulls, exchanged = 0, False, 0, 0, 0
        list_is_current = True  # fresh_list holds this rank's part of the live frontier
        lo_w, hi_w = self.rank * self.wper, (self.rank + 1) * self.wper
        next_slice, counts = self.send[: self.wper], self.send[self.wper:]
        while n_f > 0:
            level += 1
            if not pulling:
                if m_f > m_u / self.alpha and n_f > prev_n_f:
                    pulling = True
            elif n_f < self.n_global / self.beta and n_f < prev_n_f:
                pulling = False
            t0 = _time.perf_counter()
            counts.zero_()
            if pulling:
                # owner-only level: the single-GPU pull kernel on the owned rows writes depth, visited and the
                # next slice itself; nothing to exchange before the all_gather
                pulls += 1
                self.backend.pull(level, self.frontier_bits, self.visited_bits, next_slice, self.depth_local, counts)
                list_is_current = False
                t0 = mark("pull", t0)
            else:
                if not list_is_current:  # previous level was bottom-up: rebuild this rank's sparse frontier
                    self.backend.gather_fresh(self.frontier_bits[lo_w:hi_w], self.fresh_list)
                self.backend.step(False, self.frontier_bits, self.visited_bits, self.candidate_bits, self.fresh_list,
                                  my_count)
                list_is_current = True
                t0 = mark("push", t0)
                # candidate slices go to their owners; the owner ORs the P contributions inside absorb
                dist.all_to_all_single(self.a2a_recv, self.candidate_bits[: self.words])
                exchanged += (self.world - 1) * self.wper * 4
                t0 = mark("all_to_all", t0)
                self.backend.absorb(level, self.a2a_recv, self.world, self.wper, self.visited_bits, next_slice,
                                    self.depth_local, self.fresh_list, counts)
                t0 = mark("absorb", t0)
            dist.all_gather_into_tensor(self.recv, self.send)
            t0 = mark("all_gather", t0)
            exchanged += (self.world - 1) * (self.wper + 4) * 4
            # one kernel unpacks the gathered rows into the frontier bitmap, folds it into visited and collects the
            # counters; reading them is the one host sync of the level
            self.backend.merge(self.recv, self.world, self.wper, self.frontier_bits, self.visited_bits, self.counts_dev)
            self.counts_host.copy_(self.counts_dev, non_blocking=True)
            self._sync()
            per_rank = self.counts_host.view(self.world, 2)
            my_count = int(per_rank[self.rank, 0])
            prev_n_f, n_f, m_f = n_f, int(per_rank[:, 0].sum()), int(per_rank[:, 1].sum())
            m_u -= m_f
            t0 = mark("merge+sync", t0)
        self.levels, self.pull_levels, self.bytes_exchanged = level, pulls, exchanged
        return {"iterations": level, "pull_steps": pulls, "push_steps": level - pulls,
                "nvlink_bytes_received": exchanged, "enact_ms": 0.0}

    def reached_work(self):
        """(n', m') of the last BFS, summed over ranks."""
        r = self.depth_local != INF
        t = torch.stack([r.sum().to(torch.int64), self.deg_local[r].sum()])
        dist.all_reduce(t)
        return int(t[0]), int(t[1])

    def gather_depth(self) -> torch.Tensor:
        """Full depth array on every rank (tests and validation)."""
        full = torch.empty(self.n_global, dtype=torch.int32, device=self.device)
        dist.all_gather_into_tensor(full, self.depth_local)
        return full

    def pick_sources(self, count: int, seed: int = 2) -> list[int]:
        """Same seeded candidate stream as graphgen.pick_sources; owners vote on which are non-isolated."""
        from .graphgen import mix64
        pool = 64 * count + 1024
        idx = torch.arange(pool, dtype=torch.int64) + seed * 1000003
        cand = (mix64(idx) % self.n_global).to(self.device)
        mine = (cand >= self.row_begin) & (cand < self.row_begin + self.per)
        ok = torch.zeros(pool, dtype=torch.int32, device=self.device)
        ok[mine] = (self.deg_local[cand[mine] - self.row_begin] > 0).to(torch.int32)
        dist.all_reduce(ok, op=dist.ReduceOp.MAX)
        out = []
        for v, good in zip(cand.tolist(), ok.tolist()):
            if good and v not in out:
                out.append(v)
            if len(out) == count:
                break
        return out


FLT_MAX = 3.4028234663852886e38


class PartitionedSSSP:
    """SSSP over the same 1-D partition (module docstring): label-correcting rounds with a replicated array of
    tentative distances. Per round: local relax (atomic min into the replica) -> reduce_scatter(min), which
    hands every owner the best candidate of its rows ((P-1)/P * 4n bytes received per rank) -> the owner
    collects the rows that improved into its next active list; a 2-word all_reduce carries the global count.
    Same fixed point as gunrock::sssp::run (reference include/gunrock/algorithms/sssp.hxx:110-136), so the
    distances are bit-exact with the single-GPU run and the CPU oracle for non-negative weights."""

    def __init__(self, csr_local, row_begin: int, n_global: int, rank: int, world: int, backend, device):
        assert n_global % world == 0
        self.rank, self.world, self.device, self.backend = rank, world, device, backend
        self.n_global, self.per, self.row_begin = n_global, n_global // world, row_begin
        assert row_begin == rank * self.per and csr_local.offsets.numel() - 1 == self.per
        self.replica = torch.empty(n_global, dtype=torch.float32, device=device)
        self.dist_local = torch.empty(self.per, dtype=torch.float32, device=device)
        self.active = torch.zeros(self.per, dtype=torch.int32, device=device)
        self.counts = torch.zeros(2, dtype=torch.int64, device=device)
        self.deg_local = (csr_local.offsets[1:] - csr_local.offsets[:-1]).to(torch.int64)
        self.rounds = 0
        self.bytes_exchanged = 0
        self.relaxed = 0

    def sssp(self, source: int) -> dict:
        lo = self.row_begin
        owned = self.replica[lo: lo + self.per]
        self.replica.fill_(FLT_MAX)
        self.dist_local.fill_(FLT_MAX)
        self.replica[source] = 0.0
        my_count = 0
        if lo <= source < lo + self.per:
            self.dist_local[source - lo] = 0.0
            self.active[0] = source - lo
            my_count = 1
        total, rounds, exchanged, relaxed = 1, 0, 0, 0
        dense_reduce = dist.get_backend() == "gloo"  # gloo has no reduce_scatter: the CPU tests all_reduce
        while total > 0:
            rounds += 1
            self.backend.relax(self.active, my_count, self.dist_local, self.replica)
            if dense_reduce:
                dist.all_reduce(self.replica, op=dist.ReduceOp.MIN)
            else:
                dist.reduce_scatter_tensor(owned, self.replica, op=dist.ReduceOp.MIN)
            exchanged += (self.world - 1) * self.per * 4
            self.counts.zero_()
            self.backend.collect(owned, self.dist_local, self.active, self.counts)
            both = torch.cat([self.counts, self.counts])
            dist.all_reduce(both[2:])
            my_count, _, total, edges = (int(x) for x in both.tolist())  # the one host sync of the round
            relaxed += edges
        self.rounds, self.bytes_exchanged, self.relaxed = rounds, exchanged, relaxed
        return {"iterations": rounds, "nvlink_bytes_received": exchanged, "relaxed_edges": relaxed, "enact_ms": 0.0}

    def gather_dist(self) -> torch.Tensor:
        full = torch.empty(self.n_global, dtype=torch.float32, device=self.device)
        dist.all_gather_into_tensor(full, self.dist_local)
        return full

    def reached_work(self):
        r = self.dist_local != FLT_MAX
        t = torch.stack([r.sum().to(torch.int64), self.deg_local[r].sum()])
        dist.all_reduce(t)
        return int(t[0]), int(t[1])


class NativePartitionedBFS:
    """The same partitioned BFS with the whole level loop in C++ and NCCL called directly (ess_dist_bfs):
    no Python or torch dispatch on the per-level critical path. Mirrors PartitionedBFS's interface."""

    def __init__(self, csr_local, row_begin: int, n_global: int, rank: int, world: int, device, stream=None):
        import essentials_b200 as ess
        self.ess, self.rank, self.world, self.device = ess, rank, world, device
        self.n_global, self.per, self.row_begin = n_global, n_global // world, row_begin
        self.csr = csr_local
        self.offset_bits = 64 if csr_local.offsets.element_size() == 8 else 32
        self.backend = CudaBackend(csr_local, row_begin, n_global, device,
                                   stream if stream is not None else torch.cuda.current_stream(device))
        self.deg_local = (csr_local.offsets[1:] - csr_local.offsets[:-1]).to(torch.int64)
        m = torch.tensor([int(csr_local.indices.numel())], dtype=torch.int64, device=device)
        dist.all_reduce(m)
        self.m_global = int(m.item())
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = (ctypes.c_ubyte * 128)()
            ess._check(ess.lib().ess_nccl_unique_id(buf), "ess_nccl_unique_id")
            uid = torch.tensor(list(buf), dtype=torch.uint8)
        uid = uid.to(device)
        dist.broadcast(uid, 0)
        raw = bytes(uid.cpu().tolist())
        h = ctypes.c_void_p()
        ess._check(ess.lib().ess_dist_create(self.backend.ctx.handle, self.backend.graph.handle, rank, world, n_global,
                                             raw, byref(h)), "ess_dist_create")
        self.handle = h
        ptr, cnt = ctypes.c_void_p(), c_int64(0)
        ess._check(ess.lib().ess_dist_depth_local(h, byref(ptr), byref(cnt)), "ess_dist_depth_local")
        self._depth_ptr, self._depth_count = ptr.value, cnt.value
        self.depth_local = torch.empty(self.per, dtype=torch.int32, device=device)
        self.levels = self.pull_levels = 0
        self.bytes_exchanged = 0

    def bfs(self, source: int, trace=None) -> dict:
        ess = self.ess
        info = ess.RunInfo()
        ess._check(ess.lib().ess_dist_bfs(self.handle, int(source), 0.0, 0.0, byref(info)), "ess_dist_bfs")
        self.levels, self.pull_levels = int(info.iterations), int(info.pull_steps)
        self.bytes_exchanged = int(info.reserved[0])
        return {"iterations": self.levels, "pull_steps": self.pull_levels, "push_steps": int(info.push_steps),
                "nvlink_bytes_received": self.bytes_exchanged, "enact_ms": float(info.enact_ms)}

    def sssp(self, source: int) -> dict:
        """Partitioned SSSP with the round loop in C++ (ess_dist_sssp); distances via gather_dist()."""
        ess = self.ess
        info = ess.RunInfo()
        ess._check(ess.lib().ess_dist_sssp(self.handle, int(source), byref(info)), "ess_dist_sssp")
        return {"iterations": int(info.iterations), "nvlink_bytes_received": int(info.reserved[0]),
                "relaxed_edges": int(info.reserved[1]), "enact_ms": float(info.enact_ms),
                "exchange": "peer-memory" if int(info.reserved[2]) else "nccl"}

    def bfs_enactor(self, source: int, lb: str = "merge_path"):
        """gunrock::bfs::run through the enactor contract with the partitioned context (ess_dist_bfs_enactor):
        advance::execute<lb> + operators::exchange::execute per level. Returns (owned depth slice, info)."""
        ess = self.ess
        if getattr(self, "_labels_i32", None) is None:
            self._labels_i32 = torch.empty(self.n_global, dtype=torch.int32, device=self.device)
        info = ess.RunInfo()
        ess._check(ess.lib().ess_dist_bfs_enactor(self.handle, int(source), ess.LOAD_BALANCE[lb],
                                                  ess._p(self._labels_i32), byref(info)), "ess_dist_bfs_enactor")
        lo = self.row_begin
        return self._labels_i32[lo:lo + self.per], {"iterations": int(info.iterations), "enact_ms": float(info.enact_ms)}

    def sssp_enactor(self, source: int, lb: str = "merge_path"):
        """gunrock::sssp::run through the enactor contract with the partitioned context (ess_dist_sssp_enactor)."""
        ess = self.ess
        if getattr(self, "_labels_f32", None) is None:
            self._labels_f32 = torch.empty(self.n_global, dtype=torch.float32, device=self.device)
        info = ess.RunInfo()
        ess._check(ess.lib().ess_dist_sssp_enactor(self.handle, int(source), ess.LOAD_BALANCE[lb],
                                                   ess._p(self._labels_f32), byref(info)), "ess_dist_sssp_enactor")
        lo = self.row_begin
        return self._labels_f32[lo:lo + self.per], {"iterations": int(info.iterations), "enact_ms": float(info.enact_ms)}

    def _fetch_dist(self):
        if getattr(self, "dist_local", None) is None:
            self.dist_local = torch.empty(self.per, dtype=torch.float32, device=self.device)
        self.ess._check(self.ess.lib().ess_dist_copy_dist(self.handle, self.ess._p(self.dist_local)),
                        "ess_dist_copy_dist")

    def gather_dist(self) -> torch.Tensor:
        self._fetch_dist()
        full = torch.empty(self.n_global, dtype=torch.float32, device=self.device)
        dist.all_gather_into_tensor(full, self.dist_local)
        return full

    def reached_work_sssp(self):
        self._fetch_dist()
        r = self.dist_local != FLT_MAX
        t = torch.stack([r.sum().to(torch.int64), self.deg_local[r].sum()])
        dist.all_reduce(t)
        return int(t[0]), int(t[1])

    def exchange_kind(self) -> str:
        """'peer-memory' (own NVLink store kernels + epoch flags) or 'nccl' — what ess_dist_bfs uses right now."""
        kind = ctypes.c_int(0)
        self.ess._check(self.ess.lib().ess_dist_exchange_kind(self.handle, byref(kind)), "ess_dist_exchange_kind")
        return "peer-memory" if kind.value else "nccl"

    def _fetch_depth(self):
        """Copy the library-owned depth slice into self.depth_local (device to device, on the context's stream)."""
        self.ess._check(self.ess.lib().ess_dist_copy_depth(self.handle, self.ess._p(self.depth_local)),
                        "ess_dist_copy_depth")

    def reached_work(self):
        self._fetch_depth()
        r = self.depth_local != INF
        t = torch.stack([r.sum().to(torch.int64), self.deg_local[r].sum()])
        dist.all_reduce(t)
        return int(t[0]), int(t[1])

    def gather_depth(self) -> torch.Tensor:
        self._fetch_depth()
        full = torch.empty(self.n_global, dtype=torch.int32, device=self.device)
        dist.all_gather_into_tensor(full, self.depth_local)
        return full

    pick_sources = PartitionedBFS.pick_sources

    def close(self):
        if getattr(self, "handle", None):
            self.ess.lib().ess_dist_destroy(self.handle)
            self.handle = None


def _bit(b: int) -> int:
    """int32 value with only bit b set (bit 31 is the sign bit)."""
    return -(1 << 31) if b == 31 else (1 << b)


def _pack_bits(flags: torch.Tensor) -> torch.Tensor:
    """bool[n] (n % 32 == 0) -> int32[n/32], bit i of word w = flags[32w+i]."""
    f = flags.view(-1, 32).to(torch.int64)
    weights = (torch.ones(32, dtype=torch.int64, device=flags.device) << torch.arange(32, device=flags.device))
    word = (f * weights).sum(1)
    word = torch.where(word >= (1 << 31), word - (1 << 32), word)
    return word.to(torch.int32)


def build_partitioned(scale: int, edge_factor: int, rank: int, world: int, device, stream=None, seed: int = 1,
                      native: bool = True, weights: str = "none"):
    """Product constructor: regenerate the counter-based Kronecker edge list, keep this rank's rows, bind the
    CUDA backend. native=True drives the level loop from C++ with NCCL (ess_dist_bfs); False keeps the
    torch.distributed loop of PartitionedBFS."""
    from . import graphgen as gg
    n = 1 << scale
    per = n // world
    ctxmgr = torch.cuda.stream(stream) if stream is not None else _null()
    with ctxmgr:
        csr = gg.rmat_csr(scale, edge_factor, seed=seed, device=device, row_range=(rank * per, (rank + 1) * per),
                          weights=weights)
        if native:
            return NativePartitionedBFS(csr, rank * per, n, rank, world, device, stream)
        backend = CudaBackend(csr, rank * per, n, device, stream if stream is not None else torch.cuda.current_stream())
        return PartitionedBFS(csr, rank * per, n, rank, world, backend, device)


class _null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False
