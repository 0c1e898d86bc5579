"""CPU checker for the frontier-operator hot path — TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package. ``essentials_b200`` never does: the product has no CPU path.

Two libraries, both driven through ctypes on numpy arrays:

* ``liboracle.so``  — our restatement (``oracle.cpp``), always buildable (g++ only).
* ``_ref/libref_cpu.so`` — the reference's own ``*_cpu.hxx`` compiled from ``/root/reference`` by
  ``oracle/Makefile`` (int32 offsets only, single threaded).  Prebuilt here and shipped to the GPU box
  as a binary; ``/root/reference`` itself is never read at test/bench run time.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import c_float, c_int, c_int32, c_int64, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_SO = os.path.join(_HERE, "liboracle.so")
_REF_SO = os.path.join(_HERE, "_ref", "libref_cpu.so")
_REF_GPU_SO = os.path.join(_HERE, "_ref", "libref_gpu.so")


def build(verbose: bool = False) -> None:
    """(Re)build the checker libraries; the reference build is skipped when /root/reference is absent."""
    out = None if verbose else subprocess.DEVNULL
    subprocess.check_call(["make", "-s", "-f", os.path.join(_HERE, "Makefile"), "all"], stdout=out)


def _ptr(a: np.ndarray | None):
    return None if a is None else a.ctypes.data_as(c_void_p)


def _i64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.int64)


def _i32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.int32)


def _f32(a) -> np.ndarray | None:
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


_lib = None
_ref = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "oracle.cpp")
        if not os.path.exists(_ORACLE_SO) or os.path.getmtime(_ORACLE_SO) < os.path.getmtime(src):
            subprocess.check_call(["make", "-s", "-f", os.path.join(_HERE, "Makefile"), _ORACLE_SO])
        L = ctypes.CDLL(_ORACLE_SO)
        L.oracle_bfs.restype = c_float
        L.oracle_bfs.argtypes = [c_int64, c_void_p, c_void_p, c_int32, c_void_p]
        L.oracle_sssp.restype = c_float
        L.oracle_sssp.argtypes = [c_int64, c_void_p, c_void_p, c_void_p, c_int32, c_void_p]
        L.oracle_pr.restype = c_int
        L.oracle_pr.argtypes = [c_int64, c_void_p, c_void_p, c_void_p, c_float, c_float, c_int, c_int,
                                c_void_p, c_void_p]
        L.oracle_ppr.restype = c_float
        L.oracle_ppr.argtypes = [c_int64, c_void_p, c_void_p, c_int32, c_float, c_float, c_void_p]
        L.oracle_kcore.restype = c_float
        L.oracle_kcore.argtypes = [c_int64, c_void_p, c_void_p, c_void_p]
        L.oracle_randoms.restype = None
        L.oracle_randoms.argtypes = [c_int64, c_float, c_float, c_void_p]
        L.oracle_color_jacobi.restype = c_int
        L.oracle_color_jacobi.argtypes = [c_int64, c_void_p, c_void_p, c_void_p, c_void_p]
        L.oracle_color_errors.restype = c_int64
        L.oracle_color_errors.argtypes = [c_int64, c_void_p, c_void_p, c_void_p]
        L.oracle_reached_i32.restype = None
        L.oracle_reached_i32.argtypes = [c_int64, c_void_p, c_void_p, c_void_p, c_void_p]
        _lib = L
    return _lib


def have_ref() -> bool:
    return os.path.exists(_REF_SO)


def ref() -> ctypes.CDLL:
    """The reference's own CPU code (oracle/_ref). Raises if it was never built."""
    global _ref
    if _ref is None:
        if not have_ref():
            raise RuntimeError("oracle/_ref/libref_cpu.so missing: run `make -C oracle` where /root/reference exists")
        R = ctypes.CDLL(_REF_SO)
        for name in ("ref_bfs", "ref_sssp", "ref_kcore", "ref_ppr", "ref_color"):
            getattr(R, name).restype = c_float
        R.ref_bfs.argtypes = [c_int, c_int, c_void_p, c_void_p, c_int, c_void_p]
        R.ref_sssp.argtypes = [c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p]
        R.ref_kcore.argtypes = [c_int, c_int, c_void_p, c_void_p, c_void_p]
        R.ref_ppr.argtypes = [c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_float, c_float]
        R.ref_color.argtypes = [c_int, c_int, c_void_p, c_void_p, c_void_p]
        R.ref_randoms.restype = None
        R.ref_randoms.argtypes = [c_int, c_float, c_float, c_void_p]
        R.ref_load_mtx.restype = c_int
        R.ref_free.argtypes = [c_void_p]
        _ref = R
    return _ref


# ----------------------------------------------------------------------------- our restatement
def bfs(off, col, src: int, return_ms: bool = False):
    off, col = _i64(off), _i32(col)
    n = off.size - 1
    depth = np.empty(n, np.int32)
    ms = lib().oracle_bfs(n, _ptr(off), _ptr(col), int(src), _ptr(depth))
    return (depth, ms) if return_ms else depth


def sssp(off, col, w, src: int, return_ms: bool = False):
    off, col, w = _i64(off), _i32(col), _f32(w)
    n = off.size - 1
    dist = np.empty(n, np.float32)
    ms = lib().oracle_sssp(n, _ptr(off), _ptr(col), _ptr(w), int(src), _ptr(dist))
    return (dist, ms) if return_ms else dist


def pagerank(off, col, w=None, alpha=0.85, tol=1e-6, force_iters=0, max_iters=1000, return_ms=False):
    off, col, w = _i64(off), _i32(col), _f32(w)
    n = off.size - 1
    p = np.empty(n, np.float32)
    ms = c_float(0)
    it = lib().oracle_pr(n, _ptr(off), _ptr(col), _ptr(w), alpha, tol, force_iters, max_iters, _ptr(p),
                         ctypes.addressof(ms))
    return (p, it, ms.value) if return_ms else (p, it)


def ppr(off, col, seed: int, alpha=0.15, eps=1e-6):
    off, col = _i64(off), _i32(col)
    n = off.size - 1
    p = np.zeros(n, np.float32)
    lib().oracle_ppr(n, _ptr(off), _ptr(col), int(seed), alpha, eps, _ptr(p))
    return p


def kcore(off, col):
    off, col = _i64(off), _i32(col)
    n = off.size - 1
    core = np.empty(n, np.int32)
    lib().oracle_kcore(n, _ptr(off), _ptr(col), _ptr(core))
    return core


def randoms(n: int, lo: float, hi: float):
    out = np.empty(n, np.float32)
    lib().oracle_randoms(n, lo, hi, _ptr(out))
    return out


def color_jacobi(off, col, rnd=None):
    off, col = _i64(off), _i32(col)
    n = off.size - 1
    rnd = randoms(n, 0.0, float(n)) if rnd is None else _f32(rnd)
    color = np.empty(n, np.int32)
    it = lib().oracle_color_jacobi(n, _ptr(off), _ptr(col), _ptr(rnd), _ptr(color))
    return color, it


def color_errors(off, col, color) -> int:
    off, col, color = _i64(off), _i32(col), _i32(color)
    return int(lib().oracle_color_errors(off.size - 1, _ptr(off), _ptr(col), _ptr(color)))


def reached(off, depth):
    off, depth = _i64(off), _i32(depth)
    nr, mr = c_int64(0), c_int64(0)
    lib().oracle_reached_i32(off.size - 1, _ptr(off), _ptr(depth), ctypes.addressof(nr), ctypes.addressof(mr))
    return nr.value, mr.value


# ----------------------------------------------------------------------------- the reference's own CPU code
def ref_load_mtx(path: str):
    R = ref()
    n, m = c_int(), c_int()
    off, col = ctypes.POINTER(c_int)(), ctypes.POINTER(c_int)()
    val = ctypes.POINTER(c_float)()
    R.ref_load_mtx(path.encode(), ctypes.byref(n), ctypes.byref(m), ctypes.byref(off), ctypes.byref(col),
                   ctypes.byref(val))
    o = np.ctypeslib.as_array(off, (n.value + 1,)).copy()
    c = np.ctypeslib.as_array(col, (m.value,)).copy()
    v = np.ctypeslib.as_array(val, (m.value,)).copy()
    for p in (off, col, val):
        R.ref_free(ctypes.cast(p, c_void_p))
    return o, c, v


def ref_write_csr_binary(path: str, off, col, val):
    """The reference's csr_t::write_binary (formats/csr.hxx:203-234) on the given arrays."""
    off, col, val = _i32(off), _i32(col), _f32(val)
    R = ref()
    R.ref_write_csr_binary.restype = c_int
    R.ref_write_csr_binary.argtypes = [ctypes.c_char_p, c_int, c_int, c_void_p, c_void_p, c_void_p]
    R.ref_write_csr_binary(path.encode(), off.size - 1, col.size, _ptr(off), _ptr(col), _ptr(val))


def ref_read_csr_binary(path: str):
    """The reference's csr_t::read_binary (formats/csr.hxx:159-201)."""
    R = ref()
    n, m = c_int(), c_int()
    off, col = ctypes.POINTER(c_int)(), ctypes.POINTER(c_int)()
    val = ctypes.POINTER(c_float)()
    R.ref_read_csr_binary.restype = c_int
    R.ref_read_csr_binary(path.encode(), ctypes.byref(n), ctypes.byref(m), ctypes.byref(off), ctypes.byref(col),
                          ctypes.byref(val))
    o = np.ctypeslib.as_array(off, (n.value + 1,)).copy()
    c = np.ctypeslib.as_array(col, (m.value,)).copy() if m.value else np.zeros(0, np.int32)
    v = np.ctypeslib.as_array(val, (m.value,)).copy() if m.value else np.zeros(0, np.float32)
    for p in (off, col, val):
        R.ref_free(ctypes.cast(p, c_void_p))
    return o, c, v


def _ref_csr(off, col):
    off, col = _i32(off), _i32(col)
    return off, col, off.size - 1, col.size


def ref_bfs(off, col, src: int, return_ms=False):
    off, col, n, m = _ref_csr(off, col)
    d = np.empty(n, np.int32)
    ms = ref().ref_bfs(n, m, _ptr(off), _ptr(col), int(src), _ptr(d))
    return (d, ms) if return_ms else d


class RefGraph:
    """The reference's csr_t built once; bfs()/sssp() call the reference's bfs_cpu::run / sssp_cpu::run on it.
    The calls release the GIL (ctypes), so a thread pool runs several sources at once (bench.py's reference arm)."""

    def __init__(self, off, col, w=None):
        self.off, self.col, self.n, self.m = _ref_csr(off, col)
        self.w = None if w is None else _f32(w)
        R = ref()
        R.ref_graph_create.restype = c_void_p
        R.ref_graph_create.argtypes = [c_int, c_int, c_void_p, c_void_p, c_void_p]
        R.ref_graph_destroy.argtypes = [c_void_p]
        R.ref_bfs_on.restype = c_float
        R.ref_bfs_on.argtypes = [c_void_p, c_int, c_void_p]
        R.ref_sssp_on.restype = c_float
        R.ref_sssp_on.argtypes = [c_void_p, c_int, c_void_p]
        self.handle = R.ref_graph_create(self.n, self.m, _ptr(self.off), _ptr(self.col),
                                         None if self.w is None else _ptr(self.w))

    def bfs(self, src: int):
        d = np.empty(self.n, np.int32)
        ms = ref().ref_bfs_on(self.handle, int(src), _ptr(d))
        return d, float(ms)

    def sssp(self, src: int):
        d = np.empty(self.n, np.float32)
        ms = ref().ref_sssp_on(self.handle, int(src), _ptr(d))
        return d, float(ms)

    def close(self):
        if self.handle:
            ref().ref_graph_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def ref_sssp(off, col, w, src: int, return_ms=False):
    off, col, n, m = _ref_csr(off, col)
    w = _f32(w)
    d = np.empty(n, np.float32)
    ms = ref().ref_sssp(n, m, _ptr(off), _ptr(col), _ptr(w), int(src), _ptr(d))
    return (d, ms) if return_ms else d


def ref_kcore(off, col):
    off, col, n, m = _ref_csr(off, col)
    k = np.empty(n, np.int32)
    ref().ref_kcore(n, m, _ptr(off), _ptr(col), _ptr(k))
    return k


def ref_ppr(off, col, n_seeds: int, alpha=0.15, eps=1e-6):
    off, col, n, m = _ref_csr(off, col)
    p = np.zeros((n_seeds, n), np.float32)
    ref().ref_ppr(n, m, _ptr(off), _ptr(col), int(n_seeds), _ptr(p), alpha, eps)
    return p


def ref_color(off, col):
    off, col, n, m = _ref_csr(off, col)
    c = np.empty(n, np.int32)
    ref().ref_color(n, m, _ptr(off), _ptr(col), _ptr(c))
    return c


def ref_randoms(n: int, lo: float, hi: float):
    out = np.empty(n, np.float32)
    ref().ref_randoms(n, lo, hi, _ptr(out))
    return out


# ----------------------------------------------------------------------------- the reference's own GPU code
_ref_gpu = None


def have_ref_gpu() -> bool:
    return os.path.exists(_REF_GPU_SO)


def ref_gpu() -> ctypes.CDLL:
    """The reference's GPU path (block_mapped advance + Thrust filters) built for sm_100 (oracle/_ref)."""
    global _ref_gpu
    if _ref_gpu is None:
        if not have_ref_gpu():
            raise RuntimeError("oracle/_ref/libref_gpu.so missing: run `make -C oracle refgpu` where /root/reference exists")
        R = ctypes.CDLL(_REF_GPU_SO)
        for name, args in _REF_GPU_PROTOS.items():
            fn = getattr(R, name)
            fn.restype, fn.argtypes = c_float, args
        _ref_gpu = R
    return _ref_gpu


_G = [c_int, c_int, c_void_p, c_void_p, c_void_p]
_REF_GPU_PROTOS = {
    "ref_gpu_bfs": _G + [c_int, c_void_p],
    "ref_gpu_sssp": _G + [c_int, c_void_p],
    "ref_gpu_pr": _G + [c_float, c_float, c_void_p],
    "ref_gpu_ppr": _G + [c_int, c_float, c_float, c_void_p],
    "ref_gpu_kcore": _G + [c_void_p],
    "ref_gpu_color": _G + [c_void_p],
}


class _Renamed:
    """Lets ref_gpu_run call refours_* through the ref_gpu_* names."""

    def __init__(self, lib):
        self._lib = lib

    def __getattr__(self, name):
        return getattr(self._lib, name.replace("ref_gpu_", "refours_"))


def _gpu_args(csr):
    """(n, m, offsets, indices, values) device pointers of an int32 CSR held in torch CUDA tensors."""
    import torch
    assert csr.offsets.dtype == torch.int32 and csr.offsets.is_cuda, "the reference drivers are int32-only"
    vals = csr.values if csr.values is not None else torch.ones(csr.m, dtype=torch.float32, device=csr.indices.device)
    csr._ref_vals = vals  # keep alive
    return (int(csr.offsets.numel() - 1), int(csr.indices.numel()), c_void_p(csr.offsets.data_ptr()),
            c_void_p(csr.indices.data_ptr()), c_void_p(vals.data_ptr()))


_REF_ON_OURS_SO = os.path.join(_HERE, "_ref", "libref_algos_on_ours.so")
_ref_on_ours = None


def have_ref_on_ours() -> bool:
    return os.path.exists(_REF_ON_OURS_SO)


def ref_on_ours():
    """The reference's unmodified algorithm headers compiled against OUR operator headers (oracle/_ref)."""
    global _ref_on_ours
    if _ref_on_ours is None:
        R = ctypes.CDLL(_REF_ON_OURS_SO)
        proto = ref_gpu.__globals__["_REF_GPU_PROTOS"]
        for name, args in proto.items():
            fn = getattr(R, name.replace("ref_gpu_", "refours_"))
            fn.restype, fn.argtypes = c_float, args
        _ref_on_ours = R
    return _ref_on_ours


def ref_on_ours_extra(alg: str, csr, *tensors_and_params):
    """bc / spmv / hits / mst of the reference's headers on our operators (no all-reference counterpart links:
    moderngpu)."""
    import torch
    n, m, off, col, val = _gpu_args(csr)
    R = ref_on_ours()
    dev = csr.indices.device
    torch.cuda.synchronize()
    if alg == "bc":
        out = torch.zeros(n, dtype=torch.float32, device=dev)
        R.refours_bc.restype = c_float
        R.refours_bc.argtypes = _G + [c_int, c_void_p]
        ms = R.refours_bc(n, m, off, col, val, int(tensors_and_params[0]), c_void_p(out.data_ptr()))
    elif alg == "spmv":
        x = tensors_and_params[0]
        out = torch.zeros(n, dtype=torch.float32, device=dev)
        R.refours_spmv.restype = c_float
        R.refours_spmv.argtypes = _G + [c_void_p, c_void_p]
        ms = R.refours_spmv(n, m, off, col, val, c_void_p(x.data_ptr()), c_void_p(out.data_ptr()))
    elif alg == "hits":
        out = torch.zeros(1, dtype=torch.float32, device=dev)
        R.refours_hits.restype = c_float
        R.refours_hits.argtypes = _G + [c_int]
        ms = R.refours_hits(n, m, off, col, val, int(tensors_and_params[0]))
    elif alg == "mst":
        out = torch.zeros(1, dtype=torch.float32, device=dev)
        R.refours_mst.restype = c_float
        R.refours_mst.argtypes = _G + [c_void_p]
        ms = R.refours_mst(n, m, off, col, val, c_void_p(out.data_ptr()))
    elif alg == "tc":
        counts = torch.zeros(n, dtype=torch.int32, device=dev)
        total = ctypes.c_ulonglong(0)
        R.refours_tc.restype = c_float
        R.refours_tc.argtypes = _G + [c_void_p, ctypes.POINTER(ctypes.c_ulonglong)]
        ms = R.refours_tc(n, m, off, col, val, c_void_p(counts.data_ptr()), ctypes.byref(total))
        out = (counts, int(total.value))
    else:
        raise ValueError(alg)
    torch.cuda.synchronize()
    if ms < 0:
        raise RuntimeError(f"reference-on-ours {alg} failed")
    return out, float(ms)


def ref_gpu_run(alg: str, csr, *params, on_ours: bool = False):
    """Runs gunrock::<alg>::run of the REFERENCE on the GPU (on_ours=True: the reference's algorithm headers on
    our operators). Returns (result tensor, enact ms)."""
    import torch
    n, m, off, col, val = _gpu_args(csr)
    dev = csr.indices.device
    torch.cuda.synchronize()
    R = _Renamed(ref_on_ours()) if on_ours else ref_gpu()
    if alg == "bfs":
        out = torch.empty(n, dtype=torch.int32, device=dev)
        ms = R.ref_gpu_bfs(n, m, off, col, val, int(params[0]), c_void_p(out.data_ptr()))
    elif alg == "sssp":
        out = torch.empty(n, dtype=torch.float32, device=dev)
        ms = R.ref_gpu_sssp(n, m, off, col, val, int(params[0]), c_void_p(out.data_ptr()))
    elif alg == "pr":
        out = torch.empty(n, dtype=torch.float32, device=dev)
        ms = R.ref_gpu_pr(n, m, off, col, val, float(params[0]), float(params[1]), c_void_p(out.data_ptr()))
    elif alg == "ppr":
        out = torch.empty(n, dtype=torch.float32, device=dev)
        ms = R.ref_gpu_ppr(n, m, off, col, val, int(params[0]), float(params[1]), float(params[2]),
                           c_void_p(out.data_ptr()))
    elif alg == "kcore":
        out = torch.empty(n, dtype=torch.int32, device=dev)
        ms = R.ref_gpu_kcore(n, m, off, col, val, c_void_p(out.data_ptr()))
    elif alg == "color":
        out = torch.empty(n, dtype=torch.int32, device=dev)
        ms = R.ref_gpu_color(n, m, off, col, val, c_void_p(out.data_ptr()))
    else:
        raise ValueError(alg)
    torch.cuda.synchronize()
    if ms < 0:
        raise RuntimeError(f"reference GPU {alg} failed")
    return out, float(ms)
