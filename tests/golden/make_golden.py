"""Generates tests/golden/*.npz from the REFERENCE's own CPU code (oracle/_ref/libref_cpu.so, built from
/root/reference by oracle/Makefile). Run in the build container only:  python tests/golden/make_golden.py

Each fixture holds a small CSR graph plus the outputs the reference's *_cpu.hxx produced on it; tests/
compare both our oracle restatement (CPU suite) and the CUDA path (GPU suite) against them.
 - chesapeake : datasets/chesapeake/chesapeake.mtx via the reference's MatrixMarket loader + COO->CSR
                (the graph the reference CI runs, .github/workflows/ubuntu.yml:79; vectors in SURVEY.md §4)
 - rmat_s10   : essentials_b200.graphgen.rmat_csr(10, weights="hash")   (seeded, reproducible)
 - grid_24x17 : essentials_b200.graphgen.grid_csr(24, 17, weights="hash")
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from essentials_b200 import graphgen  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def outputs(off, col, val, sources, ppr_seeds):
    d = {}
    for s in sources:
        d[f"bfs_{s}"] = oracle.ref_bfs(off, col, s)
        d[f"sssp_{s}"] = oracle.ref_sssp(off, col, val, s)
    d["kcore"] = oracle.ref_kcore(off, col)
    p = oracle.ref_ppr(off, col, max(ppr_seeds) + 1)
    for s in ppr_seeds:
        d[f"ppr_{s}"] = p[s]
    n = off.size - 1
    d["randoms"] = oracle.ref_randoms(n, 0.0, float(n))
    d["sources"] = np.array(sources, np.int32)
    d["ppr_seeds"] = np.array(ppr_seeds, np.int32)
    return d


def main():
    off, col, val = oracle.ref_load_mtx("/root/reference/datasets/chesapeake/chesapeake.mtx")
    np.savez_compressed(os.path.join(OUT, "chesapeake.npz"), offsets=off, indices=col, values=val,
                        **outputs(off, col, val, [0, 7, 38], [0, 1, 2]))
    g = graphgen.rmat_csr(10, weights="hash")
    off, col, val = g.host()
    srcs = graphgen.pick_sources(g, 3)
    np.savez_compressed(os.path.join(OUT, "rmat_s10.npz"), offsets=off.astype(np.int32), indices=col, values=val,
                        **outputs(off.astype(np.int32), col, val, srcs, [0, 1]))
    g = graphgen.grid_csr(24, 17, weights="hash")
    off, col, val = g.host()
    np.savez_compressed(os.path.join(OUT, "grid_24x17.npz"), offsets=off.astype(np.int32), indices=col, values=val,
                        **outputs(off.astype(np.int32), col, val, [0, 203, 407], [0]))
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
