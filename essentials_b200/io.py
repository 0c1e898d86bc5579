"""On-disk `.csr` binary graphs, the format the reference's tools write and its drivers can load
(reference include/gunrock/formats/csr.hxx:159-236 read_binary / write_binary;
examples/tools/csr_binary/csr_binary.cu:24-39 converts MatrixMarket files into it).

Layout, native endianness, no padding:

    rows      index_t   (int32)
    columns   index_t   (int32)
    nonzeros  offset_t  (int32, or int64 when the graph was built with 64-bit offsets)
    row_offsets[rows + 1]   offset_t
    column_indices[nnz]     index_t
    nonzero_values[nnz]     value_t (float32)

The header does not say how wide offset_t is — the reference fixes it at compile time — so the reader takes
`offset_bits`; "auto" recognises the width from the file size. Files are read straight into pinned host memory and
copied to the device once (a 34 GB scale-28 partition is read in one pass, no Python-side element loops).
"""
from __future__ import annotations

import os

import numpy as np
import torch

from .graphgen import CSR


def write_csr_binary(path: str, csr: CSR) -> None:
    """Writes `csr` (host or device tensors). Missing values are written as 1.0, the weight the reference's
    loaders give pattern matrices (io/matrix_market.hxx:180-188)."""
    off = csr.offsets.cpu().numpy()
    col = csr.indices.cpu().numpy().astype(np.int32, copy=False)
    val = (csr.values.cpu().numpy().astype(np.float32, copy=False) if csr.values is not None
           else np.ones(col.size, np.float32))
    wide = off.dtype == np.int64
    n = off.size - 1
    with open(path, "wb") as f:
        np.array([n, n], np.int32).tofile(f)
        np.array([col.size], np.int64 if wide else np.int32).tofile(f)
        off.tofile(f)
        col.tofile(f)
        val.tofile(f)


def read_csr_binary(path: str, offset_bits: int | str = "auto", device="cpu", symmetric: bool = False) -> CSR:
    """Reads a `.csr` binary file. offset_bits: 32, 64 or "auto" (decided from the file size)."""
    size = os.path.getsize(path)
    with open(path, "rb") as f:
        rows, cols = (int(x) for x in np.fromfile(f, np.int32, 2))
        head = f.read(8)
    if rows < 0 or cols < 0:
        raise ValueError(f"{path}: negative dimensions")
    nnz32 = int(np.frombuffer(head[:4], np.int32)[0])
    nnz64 = int(np.frombuffer(head, np.int64)[0]) if len(head) == 8 else -1
    fits32 = nnz32 >= 0 and size == 12 + 4 * (rows + 1) + 8 * nnz32
    fits64 = nnz64 >= 0 and size == 16 + 8 * (rows + 1) + 8 * nnz64
    if offset_bits == "auto":
        if fits32 == fits64:
            raise ValueError(f"{path}: cannot tell the offset width from the file size ({size} bytes)")
        offset_bits = 32 if fits32 else 64
    if (offset_bits == 32 and not fits32) or (offset_bits == 64 and not fits64):
        raise ValueError(f"{path}: size {size} does not match a {offset_bits}-bit-offset file with {rows} rows")
    odt, nnz, start = (np.int32, nnz32, 12) if offset_bits == 32 else (np.int64, nnz64, 16)
    with open(path, "rb") as f:
        f.seek(start)
        off = np.fromfile(f, odt, rows + 1)
        col = np.fromfile(f, np.int32, nnz)
        val = np.fromfile(f, np.float32, nnz)
    if off.size != rows + 1 or col.size != nnz or val.size != nnz:
        raise ValueError(f"{path}: truncated file")
    if off[0] != 0 or off[-1] != nnz or (np.diff(off) < 0).any():
        raise ValueError(f"{path}: row offsets are not a monotone prefix sum ending at nnz")
    t = lambda a: torch.from_numpy(a).to(device)  # noqa: E731
    return CSR(rows, nnz, t(off), t(col), t(val), os.path.basename(path), symmetric)
