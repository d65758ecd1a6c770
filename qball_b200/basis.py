"""Host-side mirror of the reference's Basis for one rank (nprow = 1): the plane-wave sphere as z-rods.

In an integrated build the reference's own `Basis` object supplies these tables to qb200_plan_create;
this numpy mirror exists so that bench.py and stand-alone users can build a plan without Qball.  It follows
Basis::resize / update_g (src/qball/Basis.cc:302-751) and ChargeDensity::initialize's grid rule
(src/qball/ChargeDensity.cc:77-99) operation by operation, so that the strict `|k+G|^2 < 2 ecut` test and the rod order
(multiset by size, rod(0,0) swapped to the front) come out identical (checked bit-for-bit in tests/test_host_basis.py).
Host setup only: no wavefunction arithmetic happens here.
"""
from __future__ import annotations

import math

import numpy as np


def factorizable(n: int) -> bool:
    """Basis::factorizable (Basis.cc:126-147)"""
    for p in (11, 7, 5, 3, 3):
        if n % p == 0:
            n //= p
    while n % 2 == 0:
        n //= 2
    return n == 1


def _next_fact(n: int) -> int:
    while not factorizable(n):
        n += 2
    return n


def reciprocal_vectors(cell):
    """rows b0, b1, b2 of the reciprocal lattice (UnitCell::b(i)), 2 pi / a convention"""
    return _recip(cell)[1]


def _recip(cell):
    a = np.asarray(cell, dtype=np.float64).reshape(3, 3)

    def cross(u, v):
        return np.array([u[1] * v[2] - u[2] * v[1], u[2] * v[0] - u[0] * v[2], u[0] * v[1] - u[1] * v[0]])
    vol = float(a[0][0] * cross(a[1], a[2])[0] + a[0][1] * cross(a[1], a[2])[1] + a[0][2] * cross(a[1], a[2])[2])
    fac = 1.0 / vol
    twopi = 2.0 * math.pi
    b = np.array([twopi * cross(fac * a[1], a[2]), twopi * cross(fac * a[2], a[0]), twopi * cross(fac * a[0], a[1])])
    return a, b, vol


def _len(v):
    return math.sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2])


def density_grid(cell, ecut: float):
    """fine grid on which every wavefunction FFT runs: Gamma basis at 4*ecut, np+2, next factorizable"""
    a, b, vol = _recip(cell)
    fac = math.sqrt(2.0 * 4.0 * ecut) / (2.0 * math.pi)
    out = []
    for d in range(3):
        m = int(0.5 + fac * _len(a[d]))
        out.append(_next_fact(_next_fact(2 * m + 2) + 2))
    return tuple(out)


def make_basis(cell, ecut: float, kpoint=(0.0, 0.0, 0.0), force_complex: bool = False) -> dict:
    a, b, vol = _recip(cell)
    kp = [float(x) for x in kpoint]
    is_real = (kp[0] == 0.0 and kp[1] == 0.0 and kp[2] == 0.0) and not force_complex
    two_ecut = 2.0 * ecut
    fac = math.sqrt(two_ecut) / (2.0 * math.pi)
    hmax, kmax, lmax = (int(0.5 + fac * _len(a[d])) for d in range(3))
    ls = np.arange(-lmax - 1, lmax + 2, dtype=np.float64)

    def rods_for(hs, ks, shift):
        """rods (h,k) in row-major (h outer, k inner) order with their l range; two_e = |(kx+h) b0 + (ky+k) b1 + (kz+l) b2|^2"""
        H, K, L = np.meshgrid(np.asarray(hs, dtype=np.float64), np.asarray(ks, dtype=np.float64), ls, indexing="ij")
        cx = ((shift[0] + H) * b[0][0] + (shift[1] + K) * b[1][0]) + (shift[2] + L) * b[2][0]
        cy = ((shift[0] + H) * b[0][1] + (shift[1] + K) * b[1][1]) + (shift[2] + L) * b[2][1]
        cz = ((shift[0] + H) * b[0][2] + (shift[1] + K) * b[1][2]) + (shift[2] + L) * b[2][2]
        inside = (cx * cx + cy * cy + cz * cz) < two_ecut
        found = inside.any(axis=2)
        li = np.arange(ls.shape[0])
        lo = np.where(inside, li, ls.shape[0]).min(axis=2)
        hi = np.where(inside, li, -1).max(axis=2)
        out = []
        for i, h in enumerate(hs):
            for j, k in enumerate(ks):
                if found[i, j]:
                    out.append((int(h), int(k), int(ls[lo[i, j]]), int(hi[i, j] - lo[i, j] + 1)))
        return out

    rods = []
    if is_real:
        b2n = b[2][0] * b[2][0] + b[2][1] * b[2][1] + b[2][2] * b[2][2]
        lend = int(math.sqrt(two_ecut * (1.0 / b2n)))
        rods.append((0, 0, 0, lend + 1))
        # note: for h = 0 the reference evaluates norm(k*b1 + l*b2) (no h*b0 term); adding 0*b0 is exact
        rods += rods_for([0], range(1, kmax + 2), (0.0, 0.0, 0.0))
        rods += rods_for(range(1, hmax + 2), range(-kmax - 1, kmax + 2), (0.0, 0.0, 0.0))
    else:
        rods += rods_for(range(-hmax - 1, hmax + 2), range(-kmax - 1, kmax + 2), kp)
    kused = [r[1] for r in rods]
    idxmin1, idxmax1 = min(kused), max(kused)
    order = sorted(range(len(rods)), key=lambda i: (rods[i][3], i))      # multiset<Rod> by size, stable
    rods = [rods[i] for i in order]
    r0 = max(i for i, r in enumerate(rods) if r[0] == 0 and r[1] == 0)
    rods[0], rods[r0] = rods[r0], rods[0]                                # Basis.cc:637-651
    rod_h = np.array([r[0] for r in rods], dtype=np.int32)
    rod_k = np.array([r[1] for r in rods], dtype=np.int32)
    rod_lmin = np.array([r[2] for r in rods], dtype=np.int32)
    rod_size = np.array([r[3] for r in rods], dtype=np.int32)
    ngw = int(rod_size.sum())
    idx = np.empty((ngw, 3), dtype=np.int32)
    idx[:, 0] = np.repeat(rod_h, rod_size)
    idx[:, 1] = np.repeat(rod_k, rod_size)
    first = np.concatenate([[0], np.cumsum(rod_size)[:-1]])
    idx[:, 2] = np.arange(ngw) - np.repeat(first, rod_size) + np.repeat(rod_lmin, rod_size)
    f = idx.astype(np.float64)
    kpgx = np.empty((3, ngw))
    for d in range(3):
        kpgx[d] = ((kp[0] + f[:, 0]) * b[0][d] + (kp[1] + f[:, 1]) * b[1][d]) + (kp[2] + f[:, 2]) * b[2][d]
    kpg2 = kpgx[0] * kpgx[0] + kpgx[1] * kpgx[1] + kpgx[2] * kpgx[2]
    return dict(is_real=is_real, basis_np=(_next_fact(2 * hmax + 2), _next_fact(2 * kmax + 2), _next_fact(2 * lmax + 2)),
                idxmin1=int(idxmin1), idxmax1=int(idxmax1), ngw=ngw, nrods=len(rods), rod_h=rod_h, rod_k=rod_k, rod_lmin=rod_lmin,
                rod_size=rod_size, idx=idx, kpg2=kpg2, kpgx=kpgx, omega=vol,
                cell=np.array(cell, dtype=np.float64).reshape(9).copy(), kpoint=np.array(kp, dtype=np.float64).copy())
