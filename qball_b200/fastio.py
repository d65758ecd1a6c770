"""The reference's `save -fast` / `load -fast` wavefunction checkpoint (Wavefunction::write_fast / read_fast,
src/qball/Wavefunction.cc:2296-2403, 2651-2788): one raw binary file per task, `<filebase>NNNNNN`, holding for every
(spin, k-point) of that task the local coefficient block exactly as it sits in memory -- `mloc x nloc` complex128, column-major,
i.e. the block layout of the C ABI -- followed, when the run has empty states, by `eig[nst]` and `occ[nst]` doubles.
With band parallelism (nprow = 1) task r holds the state columns of GPU r, so the files are the natural on-disk form of the
per-GPU shards: `read_fast(...)[i]["c"]` can be handed to qb200_hpsi / qb200_compute_density as is.
"""
from __future__ import annotations

import numpy as np


def task_file(filebase: str, task: int = 0) -> str:
    """`filebase` + the six-digit task number (Wavefunction.cc:2334-2336)"""
    return f"{filebase}{task:06d}"


def read_fast(path: str, mloc: int, nloc: int, nst: int | None = None, nblocks: int = 1, has_empty: bool = False):
    """one task's file -> list (one entry per (spin, k-point) in file order) of dict(c=(nloc, mloc) complex128 [, eig, occ]).
    nst: number of states of the whole Slater determinant (length of eig / occ; default nloc)."""
    nst = nloc if nst is None else nst
    out = []
    with open(path, "rb") as f:
        for _ in range(nblocks):
            c = np.fromfile(f, dtype=np.complex128, count=mloc * nloc)
            if c.size != mloc * nloc:
                raise ValueError(f"{path}: short read ({c.size} of {mloc * nloc} coefficients)")
            blk = {"c": c.reshape(nloc, mloc)}
            if has_empty:                                   # Wavefunction.cc:2383-2391
                blk["eig"] = np.fromfile(f, dtype=np.float64, count=nst)
                blk["occ"] = np.fromfile(f, dtype=np.float64, count=nst)
                if blk["occ"].size != nst:
                    raise ValueError(f"{path}: short read of eig / occ")
            out.append(blk)
        if f.read(1):
            raise ValueError(f"{path}: trailing bytes (wrong mloc / nloc / nblocks / has_empty?)")
    return out


def write_fast(path: str, blocks) -> None:
    """inverse of read_fast: blocks = iterable of dict(c [, eig, occ]) (numpy arrays or CPU/CUDA torch tensors)"""
    with open(path, "wb") as f:
        for blk in blocks:
            c = blk["c"]
            if hasattr(c, "cpu"):
                c = c.cpu().numpy()
            np.ascontiguousarray(c, dtype=np.complex128).tofile(f)
            if "eig" in blk or "occ" in blk:
                np.ascontiguousarray(blk["eig"], dtype=np.float64).tofile(f)
                np.ascontiguousarray(blk["occ"], dtype=np.float64).tofile(f)
