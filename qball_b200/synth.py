"""Deterministic synthetic wavefunctions / potentials of a named system's shape: a counter-based
generator (splitmix64) so any state block or shard can be regenerated identically on any rank without storing it."""
from __future__ import annotations

import numpy as np


def splitmix_uniform(seed: int, start: int, count: int) -> np.ndarray:
    """u[i] in [0,1) for counters start..start+count-1 (splitmix64 finaliser, top 53 bits)"""
    with np.errstate(over="ignore"):
        z = (np.arange(start, start + count, dtype=np.uint64) + np.uint64(seed)) * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def synth_coefficients(kpg2, ecut, nst, ldc, is_real, seed=1, first_state=0):
    """c[n, ig] = (u1-1/2 + i(u2-1/2)) * exp(-|k+G|^2/(ecut/2)); Im c[G=0] = 0 for real bases; padding rows zero.
    (nst, ldc) C-contiguous complex128 == the reference's column-major mloc x nst block."""
    ngw = kpg2.shape[0]
    damp = np.exp(-kpg2 / (0.5 * ecut))
    c = np.zeros((nst, ldc), dtype=np.complex128)
    for n in range(nst):
        u = splitmix_uniform(seed, 2 * ngw * (first_state + n), 2 * ngw)
        c[n, :ngw] = ((u[0::2] - 0.5) + 1j * (u[1::2] - 0.5)) * damp
    if is_real:
        c[:, 0] = c[:, 0].real
    return c


def synth_potential(np0, np1, np2, seed=7):
    """smooth periodic v(r) (cosines, O(1) Ha) + 1e-3 fixed-seed noise; x fastest"""
    x = np.arange(np0) / np0
    y = np.arange(np1) / np1
    z = np.arange(np2) / np2
    v = (-0.7 * np.cos(2 * np.pi * x)[None, None, :] - 0.5 * np.cos(2 * np.pi * (y + 0.1))[None, :, None]
         - 0.3 * np.cos(4 * np.pi * (z + 0.2))[:, None, None])
    v = v + 0.25 * np.cos(2 * np.pi * (x[None, None, :] + y[None, :, None] - z[:, None, None]))
    n = np0 * np1 * np2
    return np.ascontiguousarray(v.reshape(n) + 1e-3 * (splitmix_uniform(seed, 0, n) - 0.5))


def synth_occ(nst, nocc=None):
    nocc = nst if nocc is None else nocc
    occ = np.zeros(nst)
    occ[:nocc] = 2.0
    if nocc < nst:
        occ[nocc - 1] = 1.25
    return occ


def _splitmix_uniform_torch(seed: int, start: int, count: int, device):
    """the same stream as splitmix_uniform, evaluated with torch int64 arithmetic (wrapping multiply, logical shifts
    emulated by masking) on `device`: bit-identical doubles, so blocks of any size can be synthesised in HBM directly"""
    import torch

    def s64(x):                         # uint64 constant -> the int64 with the same bits
        return x - (1 << 64) if x >= (1 << 63) else x

    def lsr(z, k):
        return torch.bitwise_and(torch.bitwise_right_shift(z, k), (1 << (64 - k)) - 1)

    z = (torch.arange(start, start + count, dtype=torch.int64, device=device) + seed) * s64(0x9E3779B97F4A7C15)
    z = torch.bitwise_xor(z, lsr(z, 30)) * s64(0xBF58476D1CE4E5B9)
    z = torch.bitwise_xor(z, lsr(z, 27)) * s64(0x94D049BB133111EB)
    z = torch.bitwise_xor(z, lsr(z, 31))
    return lsr(z, 11).to(torch.float64) * (1.0 / 9007199254740992.0)


def synth_coefficients_torch(kpg2, ecut, nst, ldc, is_real, seed=1, first_state=0, device="cpu", block=64):
    """synth_coefficients on a torch device (same values to the last bit up to the exp() of the damping factor, which is
    taken from numpy): returns an (nst, ldc) complex128 tensor.  Generated `block` states at a time (a dozen tensor ops per
    block, so that a profiler's launch list is not flooded with per-state helper kernels)."""
    import torch
    ngw = int(kpg2.shape[0])
    damp = torch.from_numpy(np.exp(-np.asarray(kpg2) / (0.5 * ecut))).to(device)
    c = torch.zeros((nst, ldc), dtype=torch.complex128, device=device)
    cr = torch.view_as_real(c)
    for n0 in range(0, nst, block):
        nb = min(block, nst - n0)
        # the counters of state n are 2*ngw*(first_state + n) + [0, 2*ngw): contiguous over a block of states
        u = _splitmix_uniform_torch(seed, 2 * ngw * (first_state + n0), 2 * ngw * nb, device).view(nb, ngw, 2)
        cr[n0:n0 + nb, :ngw, 0] = (u[:, :, 0] - 0.5) * damp
        cr[n0:n0 + nb, :ngw, 1] = (u[:, :, 1] - 0.5) * damp
        del u
    if is_real:
        cr[:, 0, 1] = 0.0
    return c
