"""CPU suite: the host-visible semantics of operators::exchange::execute (include/gunrock/framework/operators/
exchange/exchange.hxx) and of the partitioned enactor loop, executed over gloo on world sizes 2 and 4.

The CUDA operator cannot run here; this is its executable statement (as tests/test_peer_protocol_model.py is for the
peer-memory protocol): every rank holds a row range of the graph with GLOBAL column ids, frontiers hold LOCAL row ids,
operators see GLOBAL ids, label arrays are full-length, and one iteration is

    advance   (unchanged BFS / SSSP lambda on the owned rows: atomic::min on labels[neighbour], keep iff it improved)
    exchange  bin the kept neighbours by owner -> all_gather of the P counts -> all_to_all of (id, label) records ->
              owner: records of its own are already applied; every other record folds in with min and counts iff it
              lowered the label; improved vertices join the next frontier once (per-level bitmap) as LOCAL ids ->
              all_reduce of the frontier sizes; the loop ends when the frontier is empty on EVERY rank.

The owned label slices must equal the single-process oracle bit for bit (min-fixed points), for BFS and SSSP."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from essentials_b200 import graphgen as gg

INF_I = np.int32(2**31 - 1)
INF_F = np.float32(3.4028234663852886e38)


def _exchange(kept, labels, rank, world, per):
    """kept: GLOBAL ids the local advance kept (duplicates allowed). Returns (new local frontier, global size)."""
    owners = kept // per
    counts = np.bincount(owners, minlength=world)
    matrix = [None] * world
    dist.all_gather_object(matrix, counts.tolist())                      # every rank learns the P x P matrix
    send = [np.stack([kept[owners == p], labels[kept[owners == p]].view(np.int32)], 1) for p in range(world)]
    inbox = [None] * world
    dist.all_gather_object(inbox, send)                                  # all_to_all_v: take the column addressed to me
    seen, fresh = set(), []
    for p in range(world):
        rec = inbox[p][rank]
        assert len(rec) == matrix[p][rank], "counts matrix and records disagree"
        for v, bits in rec:
            improved = True
            if p != rank:                                                # own records are already in place
                value = np.array([bits], np.int32).view(labels.dtype)[0]
                improved = value < labels[v]
                if improved:
                    labels[v] = value
            if improved and v not in seen:                               # per-level bitmap: each vertex once
                seen.add(int(v))
                fresh.append(int(v) - rank * per)
    sizes = [None] * world
    dist.all_gather_object(sizes, len(fresh))
    return np.array(fresh, np.int64), int(sum(sizes))


def _run(off, col, val, labels, source, rank, world, per, bfs):
    lo = rank * per
    labels[:] = INF_I if bfs else INF_F
    labels[source] = 0
    frontier = np.array([source - lo], np.int64) if lo <= source < lo + per else np.array([], np.int64)
    sizes = [None] * world
    dist.all_gather_object(sizes, int(frontier.size))                    # enact(): global size after prepare_frontier
    total, iteration = sum(sizes), 0
    while total > 0:
        kept = []
        for v in frontier:                                               # advance on the owned rows
            for e in range(off[v], off[v + 1]):
                u = col[e]
                cand = np.int32(iteration + 1) if bfs else np.float32(labels[lo + v] + val[e])
                if cand < labels[u]:                                     # old = atomic::min(...); keep iff cand < old
                    labels[u] = cand
                    kept.append(u)
        frontier, total = _exchange(np.array(kept, np.int64), labels, rank, world, per)
        iteration += 1
    return iteration


def _worker(rank, world, port, scale, sources, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = 1 << scale
        per = n // world
        csr = gg.rmat_csr(scale, 8, row_range=(rank * per, (rank + 1) * per), weights="hash")
        off, col, val = csr.offsets.numpy().astype(np.int64), csr.indices.numpy().astype(np.int64), csr.values.numpy()
        results = {}
        for s in sources:
            depth, dists = np.empty(n, np.int32), np.empty(n, np.float32)
            it_b = _run(off, col, val, depth, s, rank, world, per, bfs=True)
            it_s = _run(off, col, val, dists, s, rank, world, per, bfs=False)
            pieces_d, pieces_f = [None] * world, [None] * world
            dist.all_gather_object(pieces_d, depth[rank * per:(rank + 1) * per])
            dist.all_gather_object(pieces_f, dists[rank * per:(rank + 1) * per])
            results[s] = (np.concatenate(pieces_d), np.concatenate(pieces_f), it_b, it_s)
        if rank == 0:
            out.put(results)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_exchange_semantics_reproduce_the_oracle(world):
    scale = 9
    full = gg.rmat_csr(scale, 8, weights="hash")
    off, col, val = full.host()
    sources = gg.pick_sources(full, 2) + [0]
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, scale, sources, out)) for r in range(world)]
    for p in procs:
        p.start()
    results = out.get()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    for s in sources:
        depth, dists, it_b, it_s = results[s]
        assert np.array_equal(depth, oracle.bfs(off, col, s)), s
        assert np.array_equal(dists, oracle.sssp(off, col, val, s)), s
        assert it_b >= 1 and it_s >= it_b - 1  # the loop ran until the frontier was empty on every rank
