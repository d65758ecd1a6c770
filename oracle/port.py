"""oracle/port.py -- TEST INFRASTRUCTURE ONLY: ctypes view of oracle/liboracle.so (the plain-C restatement).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")


def build(force: bool = False) -> str:
    src = [os.path.join(HERE, f) for f in ("qb_oracle.c", "qb_oracle.h")]
    if force or not os.path.exists(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in src):
        subprocess.run(["make", "-C", HERE, "-s", "port"], check=True)
    return LIB


class _Basis(C.Structure):
    _fields_ = [("is_real", C.c_int), ("np", C.c_int * 3), ("idxmin", C.c_int * 3), ("idxmax", C.c_int * 3),
                ("ngw", C.c_int), ("nrods", C.c_int),
                ("rod_h", C.POINTER(C.c_int)), ("rod_k", C.POINTER(C.c_int)), ("rod_lmin", C.POINTER(C.c_int)),
                ("rod_size", C.POINTER(C.c_int)), ("rod_first", C.POINTER(C.c_int)), ("idx", C.POINTER(C.c_int)),
                ("kpg2", C.POINTER(C.c_double)), ("kpgx", C.POINTER(C.c_double)), ("omega", C.c_double),
                ("b", C.c_double * 9)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        dp, ip, vp = C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_void_p
        L.qbo_factorizable.restype = C.c_int
        L.qbo_basis_create.restype = C.POINTER(_Basis)
        L.qbo_basis_create.argtypes = [dp, C.c_double, dp, C.c_int]
        L.qbo_basis_destroy.argtypes = [C.POINTER(_Basis)]
        L.qbo_density_grid.argtypes = [dp, C.c_double, ip]
        L.qbo_ft_create.restype = vp
        L.qbo_ft_create.argtypes = [C.c_int] * 4 + [ip] * 4 + [C.c_int] * 3
        L.qbo_ft_destroy.argtypes = [vp]
        L.qbo_ft_nvec.argtypes = [vp]
        L.qbo_ft_ntrans0.argtypes = [vp]
        L.qbo_backward.argtypes = [vp, dp, dp]
        L.qbo_forward.argtypes = [vp, dp, dp]
        L.qbo_backward_pair.argtypes = [vp, dp, dp, dp]
        L.qbo_forward_pair.argtypes = [vp, dp, dp, dp]
        L.qbo_rs_mul_add.argtypes = [vp, C.c_int, C.c_int, C.c_int, dp, dp, dp]
        L.qbo_compute_density.argtypes = [vp, C.c_int, C.c_int, dp, dp, dp]
        L.qbo_kinetic_add.argtypes = [C.c_int, C.c_int, C.c_int, dp, dp, dp]
        L.qbo_psda_update.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, dp, dp, dp, dp, dp, dp, C.c_int]
        L.qbo_psda_update.restype = C.c_double
        L.qbo_ekin_sums.argtypes = [C.c_int, C.c_int, C.c_int, dp, dp, C.c_double, dp, dp, dp, dp, dp, dp]
        L.qbo_nl_energy_species.restype = C.c_double
        L.qbo_nl_energy_species.argtypes = [C.c_int, C.c_int, C.c_int, dp, dp, C.c_int, C.c_int, C.c_int, ip, dp, dp, dp,
                                            dp, C.c_double, C.c_int, C.c_int, dp]
        L.qbo_compute_current.argtypes = [vp, C.c_int, C.c_int, C.c_int, dp, dp, dp, dp]
        L.qbo_residual.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, dp, dp, dp]
        L.qbo_gram.restype = C.c_int
        L.qbo_gram.argtypes = [C.c_int, C.c_int, C.c_int, dp]
        _lib = L
    return _lib


def _d(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _i(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def make_basis(cell, ecut, kpoint=(0.0, 0.0, 0.0), force_complex=False) -> dict:
    """Basis tables for one rank (same keys as refdrive.read_basis, minus the grid)."""
    L = lib()
    cell = np.ascontiguousarray(cell, dtype=np.float64)
    kp = np.ascontiguousarray(kpoint, dtype=np.float64)
    p = L.qbo_basis_create(_d(cell), float(ecut), _d(kp), int(force_complex))
    b = p.contents
    ngw, nr = b.ngw, b.nrods
    out = dict(is_real=bool(b.is_real), basis_np=tuple(b.np), idxmin1=b.idxmin[1], idxmax1=b.idxmax[1], ngw=ngw, nrods=nr,
               rod_h=np.ctypeslib.as_array(b.rod_h, (nr,)).copy(), rod_k=np.ctypeslib.as_array(b.rod_k, (nr,)).copy(),
               rod_lmin=np.ctypeslib.as_array(b.rod_lmin, (nr,)).copy(),
               rod_size=np.ctypeslib.as_array(b.rod_size, (nr,)).copy(),
               idx=np.ctypeslib.as_array(b.idx, (ngw, 3)).copy(), kpg2=np.ctypeslib.as_array(b.kpg2, (ngw,)).copy(),
               kpgx=np.ctypeslib.as_array(b.kpgx, (3, ngw)).copy(), omega=b.omega,
               cell=np.array(cell, dtype=np.float64).copy(), kpoint=np.array(kp, dtype=np.float64).copy())
    L.qbo_basis_destroy(p)
    return out


def density_grid(cell, ecut):
    g = np.zeros(3, dtype=np.int32)
    lib().qbo_density_grid(_d(np.ascontiguousarray(cell, dtype=np.float64)), float(ecut), _i(g))
    return tuple(int(x) for x in g)


class FT:
    """FourierTransform restatement bound to basis tables `b` (dict) and a grid."""

    def __init__(self, b: dict, np0: int, np1: int, np2: int):
        self.L = lib()
        self.np0, self.np1, self.np2 = np0, np1, np2
        self.N = np0 * np1 * np2
        self.ngw = b["ngw"]
        self.is_real = b["is_real"]
        keep = [np.ascontiguousarray(b[k], dtype=np.int32) for k in ("rod_h", "rod_k", "rod_lmin", "rod_size")]
        self.h = self.L.qbo_ft_create(np0, np1, np2, b["nrods"], *[_i(a) for a in keep], int(b["is_real"]),
                                      int(b["idxmin1"]), int(b["idxmax1"]))
        self.nvec = self.L.qbo_ft_nvec(self.h)
        self.ntrans0 = self.L.qbo_ft_ntrans0(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.qbo_ft_destroy(self.h)
            self.h = None

    def backward(self, c, c2=None):
        f = np.empty(self.N, dtype=np.complex128)
        c = np.ascontiguousarray(c, dtype=np.complex128)
        if c2 is None:
            self.L.qbo_backward(self.h, _d(c), _d(f))
        else:
            c2 = np.ascontiguousarray(c2, dtype=np.complex128)
            self.L.qbo_backward_pair(self.h, _d(c), _d(c2), _d(f))
        return f

    def forward(self, f, pair=False):
        f = np.array(f, dtype=np.complex128, copy=True)
        c1 = np.zeros(self.ngw, dtype=np.complex128)
        if not pair:
            self.L.qbo_forward(self.h, _d(f), _d(c1))
            return c1
        c2 = np.zeros(self.ngw, dtype=np.complex128)
        self.L.qbo_forward_pair(self.h, _d(f), _d(c1), _d(c2))
        return c1, c2

    def rs_mul_add(self, c, v, cp):
        """c, cp: (nst, ldc) complex128 C-contiguous; cp updated in place."""
        nst, ldc = c.shape
        assert cp.shape == c.shape and c.flags.c_contiguous and cp.flags.c_contiguous
        self.L.qbo_rs_mul_add(self.h, self.ngw, ldc, nst, _d(c), _d(np.ascontiguousarray(v)), _d(cp))
        return cp

    def compute_density(self, c, fac, rho):
        nst, ldc = c.shape
        fac = np.ascontiguousarray(fac, dtype=np.float64)
        self.L.qbo_compute_density(self.h, ldc, nst, _d(c), _d(fac), _d(rho))
        return rho


def compute_current(ft: "FT", c, fac, kpgx, cur=None):
    """CurrentDensity::update_current (CurrentDensity.cc:52-86): returns cur (3, N), accumulated"""
    nst, ldc = c.shape
    if cur is None:
        cur = np.zeros((3, ft.N))
    fac = np.ascontiguousarray(fac, dtype=np.float64)
    kpgx = np.ascontiguousarray(kpgx, dtype=np.float64)
    lib().qbo_compute_current(ft.h, ft.ngw, ldc, nst, _d(c), _d(fac), _d(kpgx), _d(cur))
    return cur


def kinetic_add(kpg2, c, cp):
    nst, ldc = c.shape
    lib().qbo_kinetic_add(kpg2.shape[0], ldc, nst, _d(np.ascontiguousarray(kpg2)), _d(c), _d(cp))
    return cp


def ekin_sums(kpg2, c, occ, is_real, kpgx=None, fstress=None, dfstress=None):
    """EnergyFunctional.cc:1155-1296 for one (spin, k-point): returns (tsum[14], psi2sum[ngw])"""
    nst, ldc = c.shape
    ngw = kpg2.shape[0]
    tsum, p2 = np.zeros(14), np.zeros(ngw)
    opt = lambda a: None if a is None else _d(np.ascontiguousarray(a, dtype=np.float64))  # noqa: E731
    lib().qbo_ekin_sums(ngw, ldc, nst, _d(c), _d(np.ascontiguousarray(occ, dtype=np.float64)), C.c_double(1.0 if is_real else 0.5),
                        _d(np.ascontiguousarray(kpg2)), opt(kpgx), opt(fstress), opt(dfstress), _d(p2), _d(tsum))
    return tsum, p2


def xc_lda(rho):
    rho = np.ascontiguousarray(rho, dtype=np.float64)
    e, v = np.zeros_like(rho), np.zeros_like(rho)
    lib().qbo_xc_lda(C.c_size_t(rho.size), _d(rho), _d(e), _d(v))
    return e, v


def xc_pbe(rho, grad):
    rho, grad = np.ascontiguousarray(rho, dtype=np.float64), np.ascontiguousarray(grad, dtype=np.float64)
    e, v1, v2 = np.zeros_like(rho), np.zeros_like(rho), np.zeros_like(rho)
    lib().qbo_xc_pbe(C.c_size_t(rho.size), _d(rho), _d(grad), _d(e), _d(v1), _d(v2))
    return e, v1, v2


def g2i_of(b):
    """Basis::g2i (k = 0 basis): 1/|G|^2, 0 at G = 0"""
    g2 = np.asarray(b["kpg2"])
    return np.where(g2 > 0.0, 1.0 / np.where(g2 > 0.0, g2, 1.0), 0.0)


def update_vhxc(vft, vb, xc, rhor, rhog, vion, rhopst):
    """EnergyFunctional::update_vhxc for one spin on the density basis vb (k = 0: gx = kpgx): returns (v_r, [exc, eps, ehart])"""
    N = vft.N
    v_r, en = np.zeros(N), np.zeros(3)
    cc = lambda a: np.ascontiguousarray(a, dtype=np.complex128)  # noqa: E731
    L = lib()
    L.qbo_update_vhxc.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int] + [C.POINTER(C.c_double)] * 6 + [C.c_double] + [C.POINTER(C.c_double)] * 2
    L.qbo_update_vhxc(vft.h, int(xc), int(vb["ngw"]), int(vb["is_real"]), _d(np.ascontiguousarray(rhor)), _d(cc(rhog)),
                      _d(np.ascontiguousarray(vb["kpgx"])), _d(g2i_of(vb)), _d(cc(vion)), _d(cc(rhopst)), C.c_double(vb["omega"]), _d(v_r), _d(en))
    return v_r, en


def preconditioner_diag(kpg2, ecutprec, fstress=None):
    """Preconditioner::update (Preconditioner.cc:47-90): diag[ig] = 0.5/max(e, ecutprec), e = 0.5 (|k+G|^2 [+ fstress])"""
    e = 0.5 * (kpg2 + (0.0 if fstress is None else fstress))
    return np.where(e < ecutprec, 0.5 / ecutprec, 0.5 / np.maximum(e, 1e-300))


def psda_update(ngw, is_real, c, dc, c_last, dc_last, occ, precdiag, extrapolate):
    """in place; returns theta before clipping"""
    nst, ldc = c.shape
    L = lib()
    L.qbo_psda_update.restype = C.c_double
    return float(L.qbo_psda_update(ngw, ldc, nst, int(is_real), _d(c), _d(dc), _d(c_last), _d(dc_last),
                                   _d(np.ascontiguousarray(occ, dtype=np.float64)), _d(np.ascontiguousarray(precdiag)), int(extrapolate)))


def nl_energy(b: dict, c, occ, species, compute_hpsi=True, cp=None):
    """NonLocalPotential::energy NC branch over all species; species = list of dict(na,npr,lproj,wt,twnl,tau)."""
    nst, ldc = c.shape
    if cp is None:
        cp = np.zeros_like(c)
    namax = max([s["na"] for s in species] + [0])
    nab = 128 if namax > 128 else namax  # NonLocalPotential.cc:1537-1541
    kpgx = np.ascontiguousarray(b["kpgx"], dtype=np.float64)
    occ = np.ascontiguousarray(occ, dtype=np.float64)
    enl = 0.0
    for s in species:
        if s["npr"] == 0 or s["na"] == 0:
            continue
        lproj = np.ascontiguousarray(s["lproj"], dtype=np.int32)
        wt = np.ascontiguousarray(s["wt"], dtype=np.float64)
        twnl = np.ascontiguousarray(s["twnl"], dtype=np.float64)
        tau = np.ascontiguousarray(s["tau"], dtype=np.float64)
        enl += lib().qbo_nl_energy_species(b["ngw"], ldc, nst, _d(c), _d(occ), int(b["is_real"]), s["na"], s["npr"],
                                           _i(lproj), _d(wt), _d(twnl), _d(tau), _d(kpgx), float(b["omega"]), nab,
                                           int(compute_hpsi), _d(cp))
    return enl, cp


def hpsi(b: dict, ft: FT, c, v, occ, species):
    """Whole H psi in the reference's order: clear -> nonlocal -> kinetic -> local (EnergyFunctional.cc:1142-1701)."""
    cp = np.zeros_like(c)
    enl, _ = nl_energy(b, c, occ, species, True, cp)
    kinetic_add(b["kpg2"], c, cp)
    ft.rs_mul_add(c, v, cp)
    return enl, cp


def residual(c, hc, is_real):
    """PSD/PSDA descent direction (PSDAWavefunctionStepper.cc:65-84, 264-277): returns (hc - c a, a) with a = c^H hc;
    c: (nall, ldc), hc: (nst, ldc) complex128; a: (nst, nall) (row n = column n of the reference's matrix a)."""
    nall, ldc = c.shape
    nst = hc.shape[0]
    out = np.array(hc, dtype=np.complex128, copy=True)
    a = np.zeros((nst, nall), dtype=np.float64 if is_real else np.complex128)
    lib().qbo_residual(ldc, nall, nst, int(is_real), _d(np.ascontiguousarray(c)), _d(out), _d(a))
    return out, a


def subspace_h(c, hc, is_real):
    """h = c^H (H c) as Wavefunction::diag forms it (Wavefunction.cc:1538-1539 real: gemm('t','n',2.0) + ger(-1.0); :1641
    complex: gemm('c','n',1.0)); h[m, n], numpy"""
    ngw_rows = c.shape[1]
    if is_real:
        cr = np.ascontiguousarray(c).view(np.float64).reshape(c.shape[0], 2 * ngw_rows)
        hr = np.ascontiguousarray(hc).view(np.float64).reshape(hc.shape[0], 2 * ngw_rows)
        return 2.0 * cr @ hr.T - np.outer(cr[:, 0], hr[:, 0])
    return c.conj() @ hc.T


def diag(c, hc, is_real):
    """eigenvalues of h from its lower triangle, ascending (syevd / heevd 'l', Wavefunction.cc:1604, 1682) and the
    eigenvectors z (columns); LAPACK through numpy, as the reference calls LAPACK"""
    h = subspace_h(c, hc, is_real)
    w, z = np.linalg.eigh(h, UPLO="L")
    return w, z, h


def gram(c, is_real):
    """SlaterDet::gram (SlaterDet.cc:1043-1143): Cholesky orthonormalisation, returns the new (nst, ldc) block"""
    nst, ldc = c.shape
    out = np.array(c, dtype=np.complex128, copy=True)
    info = lib().qbo_gram(ldc, nst, int(is_real), _d(out))
    assert info == 0, f"gram: leading minor {info} not positive definite"
    return out


# --------------------------------------------------------------------------------------------- ultrasoft beta.psi (row f4)
def us_anl(b: dict, s: dict):
    """betag with the structure factor: anl[ia*npr + lm, ig] = betag_lm(ig) * (-i)^l * exp(-i (k+G).tau_ia)
    (SlaterDet::calc_betag, SlaterDet.cc:2006-2127, and the phase of calc_betapsi's !highmem branch, :2222-2236)"""
    kpgx = np.asarray(b["kpgx"], dtype=np.float64)                       # (3, ngw)
    il = np.array([1.0, -1.0j, -1.0, 1.0j])[np.asarray(s["lproj"]) & 3]     # (-i)^l
    arg = np.asarray(s["tau"]) @ kpgx                                     # (na, ngw)
    ph = np.cos(arg) - 1j * np.sin(arg)
    bg = np.asarray(s["twnl"]) * il[:, None]                              # (npr, ngw)
    return (ph[:, None, :] * bg[None, :, :]).reshape(s["na"] * s["npr"], -1)


def us_qmatrix(s: dict):
    """dense symmetric coupling of one species from its (lm1, lm2, qaug) triples (the loop of SlaterDet::calc_spsi,
    SlaterDet.cc:2536-2549: bpsum[lm1] += q * bp[lm2]; and the transposed term when lm1 != lm2)"""
    q = np.zeros((s["npr"], s["npr"]))
    for a, bb, v in zip(s["lm1"], s["lm2"], s["qaug"]):
        q[a, bb] += v
        if a != bb:
            q[bb, a] += v
    return q


def us_betapsi(b: dict, c, species):
    """SlaterDet::calc_betapsi (SlaterDet.cc:2130-2263, complex basis): betapsi[n, p] = sum_G conj(anl_p(G)) c_n(G);
    species concatenated, p = ia*npr + lm inside a species"""
    ngw = b["ngw"]
    return np.concatenate([c[:, :ngw] @ np.conj(us_anl(b, s)).T for s in species], axis=1)


def us_add_beta(b: dict, f, species, cp):
    """cp_n(G) += sum_p anl_p(G) f[n, p] (the gemm of SlaterDet::calc_spsi, SlaterDet.cc:2565)"""
    ngw, off = b["ngw"], 0
    for s in species:
        m = s["na"] * s["npr"]
        cp[:, :ngw] += f[:, off:off + m] @ us_anl(b, s)
        off += m
    return cp


def us_spsi(b: dict, c, species):
    """SlaterDet::calc_spsi (SlaterDet.cc:2426-2570): S psi = psi + sum_{a,nm} beta_n q_nm <beta_m|psi> / omega"""
    bp = us_betapsi(b, c, species)
    f = np.zeros_like(bp)
    off = 0
    for s in species:
        q = us_qmatrix(s)
        m = s["na"] * s["npr"]
        blk = bp[:, off:off + m].reshape(-1, s["na"], s["npr"])
        f[:, off:off + m] = (blk @ q.T).reshape(-1, m) / b["omega"]
        off += m
    return us_add_beta(b, f, species, c.copy()), bp


def us_pair_matrix(s: dict, values):
    """dense symmetric npr x npr matrix from per-q values along the species' (lm1, lm2) pairs"""
    m = np.zeros((s["npr"], s["npr"]))
    for a, bb, v in zip(s["lm1"], s["lm2"], values):
        m[a, bb] = v
        m[bb, a] = v
    return m


def us_dmat(s: dict, vkpgx, veff):
    """D_nm^I = D_nm^0 + sum_G Re( conj(sf_I(G) Q_nm(G)) veff(G) ) on the density basis (NonLocalPotential.cc:1607-1636,
    !highmem branch; sf_I(G) = exp(-i G.tau_I), :2724-2731).  Returns (na, nq)."""
    arg = np.asarray(s["tau"]) @ np.asarray(vkpgx)                       # (na, ngv)
    sf = np.cos(arg) - 1j * np.sin(arg)
    u = np.conj(sf) * np.asarray(veff)[None, :]                           # (na, ngv)
    return np.real(u @ np.conj(np.asarray(s["qnmg"])).T) + np.asarray(s["dzero"])[None, :]


def us_energy(b: dict, c, occ, species, vkpgx, veff, compute_hpsi=True):
    """ultrasoft branch of NonLocalPotential::energy (NonLocalPotential.cc:1554-1752, no forces):
    E_nl = sum_n occ_n/omega sum_{I,q} mult_q dzero_q Re(conj(bp[I,lm1]) bp[I,lm2]),
    H psi_n += sum_{I,lm} beta^I_lm (1/omega) sum_lm' D^I[lm,lm'] bp_n[I,lm']."""
    bp = us_betapsi(b, c, species)
    occ = np.asarray(occ, dtype=np.float64)
    enl, off = 0.0, 0
    f = np.zeros_like(bp)
    for s in species:
        m = s["na"] * s["npr"]
        blk = bp[:, off:off + m].reshape(-1, s["na"], s["npr"])           # (nst, na, npr)
        d0 = us_pair_matrix(s, s["dzero"])
        enl += float(np.einsum("n,nia,ab,nib->", occ, np.conj(blk), d0, blk).real) / b["omega"]
        if compute_hpsi:
            dm = us_dmat(s, vkpgx, veff)
            for ia in range(s["na"]):
                f[:, off + ia * s["npr"]:off + (ia + 1) * s["npr"]] = blk[:, ia, :] @ us_pair_matrix(s, dm[ia]).T / b["omega"]
        off += m
    hp = us_add_beta(b, f, species, np.zeros_like(c)) if compute_hpsi else None
    return enl, hp, bp


def us_summat(s: dict, bp_species, fac):
    """summat[I, q] = sum_n fac_n mult_q conj(bp_n[I,lm1]) bp_n[I,lm2], fac_n = weight occ_n / omega (ChargeDensity.cc:352-368)"""
    blk = np.asarray(bp_species).reshape(-1, s["na"], s["npr"])
    mult = np.where(np.asarray(s["lm1"]) == np.asarray(s["lm2"]), 1.0, 2.0)
    return np.einsum("n,niq,niq->iq", np.asarray(fac, dtype=np.float64), np.conj(blk[:, :, s["lm1"]]), blk[:, :, s["lm2"]]) * mult[None, :]


def us_rhog(b: dict, c, fac, species, vkpgx):
    """rhogus(G) = sum_{I,q} Q_q(G) summat[I,q] exp(-i G.tau_I) / omega (ChargeDensity.cc:397-428 with sfactloc_ of :800-820)"""
    bp = us_betapsi(b, c, species)
    rg = np.zeros(np.asarray(vkpgx).shape[1], dtype=np.complex128)
    off = 0
    for s in species:
        m = s["na"] * s["npr"]
        sm = us_summat(s, bp[:, off:off + m], fac)                        # (na, nq)
        arg = np.asarray(s["tau"]) @ np.asarray(vkpgx)
        sf = (np.cos(arg) - 1j * np.sin(arg)) / b["omega"]
        rg += np.einsum("iq,qg,ig->g", sm, np.asarray(s["qnmg"]), sf)
        off += m
    return rg


def us_augment_density(b: dict, vft: "FT", c, fac, species, vkpgx, rho):
    """rho += Re FT^-1[rhogus] (ChargeDensity.cc:437-456); returns (rho, uscharge = sum Re(.) omega / N)"""
    r = vft.backward(us_rhog(b, c, fac, species, vkpgx)).real
    rho += r
    return rho, float(r.sum()) * b["omega"] / r.size


def update_twnl(kpgx, lproj, mproj, tabproj, gspl, vnlg, vnlg_spl, gcut=None):
    """NonLocalPotential::update_twnl for a Kleinman-Bylander species (NonLocalPotential.cc:261-1522, the twnl part, no stress
    derivatives): twnl[ipr, ig] = Y_lm(k+G) v(|k+G|), v = the species' radial cubic spline (Species::dvnlg, Species.cc:1492-1505:
    0 beyond the last knot; splintd, spline.cc:126-156), real spherical harmonics in the reference's order and normalisation
    (l=0 :334, l=1 :466-470, l=2 :712-747, l=3 :1126-1140).  kpgx: (3, ngw).  Arithmetic in the reference's order."""
    kpgx = np.asarray(kpgx, dtype=np.float64)
    x, y, z = kpgx[0], kpgx[1], kpgx[2]
    kpg = np.sqrt(x * x + y * y + z * z)                 # Basis::update_g: kpg_ = sqrt(kpg2_) (Basis.cc:733-737)
    with np.errstate(divide="ignore"):
        gi = np.where(kpg > 0.0, 1.0 / kpg, 0.0)
    gspl = np.asarray(gspl, dtype=np.float64)
    gcut = float(gspl[-1]) if gcut is None else float(gcut)
    n = gspl.shape[0]
    khi = np.clip(np.searchsorted(gspl, kpg, side="right"), 1, n - 1)   # first knot > x (the bisection of spline.cc:135-142)
    klo = khi - 1
    h = gspl[khi] - gspl[klo]
    a = (gspl[khi] - kpg) / h
    bb = (kpg - gspl[klo]) / h
    pi = np.pi
    fpi = 4.0 * pi
    s14pi, s34pi, s54pi, s3 = np.sqrt(1.0 / fpi), np.sqrt(3.0 / fpi), np.sqrt(5.0 / fpi), np.sqrt(3.0)
    s74pi, s2132pi, s3532pi, s1054pi = np.sqrt(7.0 / fpi), np.sqrt(21.0 / (32. * pi)), np.sqrt(35.0 / (32. * pi)), np.sqrt(105.0 / fpi)
    gi2 = gi * gi
    gi3 = gi2 * gi
    xx, yy, zz = x * x * gi2, y * y * gi2, z * z * gi2
    xy, yz, xz = x * y * gi2, y * z * gi2, x * z * gi2
    ylm = {
        (0, 0): s14pi,
        (1, 0): s34pi * x * gi, (1, 1): s34pi * y * gi, (1, 2): s34pi * z * gi,
        (2, 0): s54pi * 0.5 * (3.0 * zz - 1.0), (2, 1): s54pi * 0.5 * s3 * (xx - yy), (2, 2): s54pi * s3 * xy,
        (2, 3): s54pi * s3 * yz, (2, 4): s54pi * s3 * xz,
        (3, 0): s74pi * 0.5 * z * gi * (5.0 * zz - 3.0), (3, 1): s2132pi * x * gi * (5.0 * zz - 1.0),
        (3, 2): s2132pi * y * gi * (5.0 * zz - 1.0), (3, 3): s1054pi * x * y * z * gi3,
        (3, 4): s1054pi * 0.5 * z * gi * (xx - yy), (3, 5): s3532pi * x * gi * (xx - 3.0 * yy),
        (3, 6): s3532pi * y * gi * (3.0 * xx - yy),
    }
    out = np.zeros((len(lproj), kpg.shape[0]))
    vcache = {}
    for ipr, (l, m, t) in enumerate(zip(lproj, mproj, tabproj)):
        t = int(t)
        if t not in vcache:
            ya, y2a = np.asarray(vnlg[t]), np.asarray(vnlg_spl[t])
            v = a * ya[klo] + bb * ya[khi] + h * h * (1.0 / 6.0) * ((a * a * a - a) * y2a[klo] + (bb * bb * bb - bb) * y2a[khi])
            vcache[t] = np.where(kpg > gcut, 0.0, v)
        out[ipr] = ylm[(int(l), int(m))] * vcache[t]
    return out


def _ylm_table(kpgx):
    """real spherical harmonics of update_twnl (reference order and normalisation) for every plane wave: dict (l, m) -> array"""
    kpgx = np.asarray(kpgx, dtype=np.float64)
    x, y, z = kpgx[0], kpgx[1], kpgx[2]
    kpg = np.sqrt(x * x + y * y + z * z)
    with np.errstate(divide="ignore"):
        gi = np.where(kpg > 0.0, 1.0 / kpg, 0.0)
    pi = np.pi
    fpi = 4.0 * pi
    s14pi, s34pi, s54pi, s3 = np.sqrt(1.0 / fpi), np.sqrt(3.0 / fpi), np.sqrt(5.0 / fpi), np.sqrt(3.0)
    s74pi, s2132pi, s3532pi, s1054pi = np.sqrt(7.0 / fpi), np.sqrt(21.0 / (32. * pi)), np.sqrt(35.0 / (32. * pi)), np.sqrt(105.0 / fpi)
    gi2 = gi * gi
    gi3 = gi2 * gi
    xx, yy, zz = x * x * gi2, y * y * gi2, z * z * gi2
    xy, yz, xz = x * y * gi2, y * z * gi2, x * z * gi2
    return kpg, gi, {
        (0, 0): s14pi + 0.0 * kpg,
        (1, 0): s34pi * x * gi, (1, 1): s34pi * y * gi, (1, 2): s34pi * z * gi,
        (2, 0): s54pi * 0.5 * (3.0 * zz - 1.0), (2, 1): s54pi * 0.5 * s3 * (xx - yy), (2, 2): s54pi * s3 * xy,
        (2, 3): s54pi * s3 * yz, (2, 4): s54pi * s3 * xz,
        (3, 0): s74pi * 0.5 * z * gi * (5.0 * zz - 3.0), (3, 1): s2132pi * x * gi * (5.0 * zz - 1.0),
        (3, 2): s2132pi * y * gi * (5.0 * zz - 1.0), (3, 3): s1054pi * x * y * z * gi3,
        (3, 4): s1054pi * 0.5 * z * gi * (xx - yy), (3, 5): s3532pi * x * gi * (xx - 3.0 * yy),
        (3, 6): s3532pi * y * gi * (3.0 * xx - yy),
    }


def update_twnl_semilocal(kpgx, lproj, mproj, rproj):
    """NonLocalPotential::update_twnl for a semi-local species (nquad > 0; NonLocalPotential.cc:366-419 l=0, :500-600 l=1,
    :800-960 l=2, :1230-1345 l=3; no stress derivatives): twnl[ipr, ig] = Y_lm(k+G) 4 pi j_l(|k+G| r) r at the projector's
    quadrature radius r, spherical Bessel functions written with sin / cos as the reference does (l = 0: 4 pi sin(qr)/q, and
    4 pi r at q = 0; l >= 1: 0 at q r = 0)."""
    kpg, gi, ylm = _ylm_table(kpgx)
    fpi = 4.0 * np.pi
    out = np.zeros((len(lproj), kpg.shape[0]))
    for ipr, (l, m, r) in enumerate(zip(lproj, mproj, rproj)):
        l, m, r = int(l), int(m), float(r)
        zz = kpg * r
        with np.errstate(divide="ignore", invalid="ignore"):
            zi = np.where(zz != 0.0, 1.0 / zz, 0.0)
            s, c = np.sin(zz), np.cos(zz)
            if l == 0:
                v = np.where((gi == 0.0) & (kpg == 0.0), fpi * r, fpi * np.sin(kpg * r) * gi)
            elif l == 1:
                v = np.where(zz != 0.0, fpi * ((s * zi - c) * zi) * r, 0.0)
            elif l == 2:
                v = np.where(zz != 0.0, fpi * (((3.0 * zi * zi - 1.0) * s - 3.0 * zi * c) * zi) * r, 0.0)
            else:
                v = np.where(zz != 0.0, fpi * ((15.0 * zi * zi - 6.0) * zi * zi * s - (15.0 * zi * zi - 1.0) * zi * c) * r, 0.0)
        out[ipr] = ylm[(l, m)] * v
    return out
