"""Host-side mirror of the reference's classes on this path, over the C ABI (names and argument meaning follow
LLNL/qball: FourierTransform.h:144-153, SlaterDet.h:113-115, NonLocalPotential.h:93-96).

The C++ equivalent a Qball maintainer would use is include/qball_b200.hpp; this Python mirror exists so that the
parity tests and bench.py read like the reference's own drivers (src/tests/testFourierTransform.cc).
Arrays: numpy (host) or CUDA torch tensors (device, used in place).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi


class FourierTransform:
    """FourierTransform(basis, np0, np1, np2)  (FourierTransform.cc:144-526), one rank.

    `basis` is any object/dict with the reference Basis accessors' values: nrods (nrod_loc), rod_h, rod_k, rod_lmin,
    rod_size, is_real (Basis::real()), idxmin1/idxmax1 (Basis::idxmin(1)/idxmax(1))."""

    def __init__(self, basis, np0: int, np1: int, np2: int, device: int = 0, stream=None):
        L = capi.load()
        g = basis.get if isinstance(basis, dict) else (lambda k: getattr(basis, k))
        self._keep = [capi._iarr(g(k)) for k in ("rod_h", "rod_k", "rod_lmin", "rod_size")]
        h = C.c_void_p()
        capi._check(L.qb200_plan_create(C.byref(h), device, np0, np1, np2, int(g("nrods")), *[k[1] for k in self._keep],
                                        int(bool(g("is_real"))), int(g("idxmin1")), int(g("idxmax1"))), "qb200_plan_create")
        self._h, self._L = h, L
        self.np0_, self.np1_, self.np2_ = np0, np1, np2
        self.is_real = bool(g("is_real"))
        if stream is not None:
            self.set_stream(stream)

    # reference accessors (FourierTransform.h:121-139)
    def np0(self): return self.np0_
    def np1(self): return self.np1_
    def np2(self): return self.np2_
    def np012(self): return self.np0_ * self.np1_ * self.np2_
    def np012loc(self): return self.np012()
    def nvec(self): return int(self._L.qb200_plan_query(self._h, 3))
    def ntrans0(self): return int(self._L.qb200_plan_query(self._h, 4))
    def ngw(self): return int(self._L.qb200_plan_query(self._h, 5))
    def fused(self): return bool(self._L.qb200_plan_query(self._h, 7))
    def batch(self): return int(self._L.qb200_plan_query(self._h, 8))
    def launches(self): return int(self._L.qb200_plan_query(self._h, 9))
    def query(self, what): return int(self._L.qb200_plan_query(self._h, int(what)))

    def set_stream(self, stream):
        s = stream.cuda_stream if hasattr(stream, "cuda_stream") else int(stream)
        capi._check(self._L.qb200_plan_set_stream(self._h, s), "qb200_plan_set_stream")

    def set_workspace(self, nbytes: int):
        capi._check(self._L.qb200_plan_set_workspace(self._h, int(nbytes)), "qb200_plan_set_workspace")

    def set_coefficient_tag(self, tag: int):
        """content version of the HOST coefficient blocks passed from now on (0: always upload); see qball_b200.h"""
        capi._check(self._L.qb200_plan_set_coefficient_tag(self._h, int(tag)), "qb200_plan_set_coefficient_tag")

    def backward(self, c, f, c2=None):
        """backward(c, f) / backward(c1, c2, f): f(r) = sum_G c_G e^{+iGr}  (FourierTransform.cc:529-567)"""
        if c2 is None:
            capi._check(self._L.qb200_fft_backward(self._h, capi.ptr(c), capi.ptr(f)), "qb200_fft_backward")
        else:
            capi._check(self._L.qb200_fft_backward_pair(self._h, capi.ptr(c), capi.ptr(c2), capi.ptr(f)), "qb200_fft_backward_pair")
        return f

    def forward(self, f, c, c2=None):
        """forward(f, c) / forward(f, c1, c2); f is clobbered as in the reference (FourierTransform.cc:542-581)"""
        if c2 is None:
            capi._check(self._L.qb200_fft_forward(self._h, capi.ptr(f), capi.ptr(c)), "qb200_fft_forward")
        else:
            capi._check(self._L.qb200_fft_forward_pair(self._h, capi.ptr(f), capi.ptr(c), capi.ptr(c2)), "qb200_fft_forward_pair")
        return c

    def close(self):
        if getattr(self, "_h", None):
            self._L.qb200_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _block_dims(c):
    """(nst, ldc) of a coefficient block given as a 2-D (nst, ldc) complex array/tensor (== column-major ldc x nst)"""
    assert len(c.shape) == 2
    return int(c.shape[0]), int(c.shape[1])


def rs_mul_add(ft: FourierTransform, c, v, cp, kpg2=None):
    """SlaterDet::rs_mul_add(ft, v, sdp): cp += FT[v * FT^-1 c] (SlaterDet.cc:971-1040); kpg2 fuses the kinetic term."""
    nst, ldc = _block_dims(c)
    capi._check(ft._L.qb200_rs_mul_add(ft._h, ldc, nst, capi.ptr(c), capi.ptr(v), capi.ptr(kpg2), capi.ptr(cp)), "qb200_rs_mul_add")
    return cp


def compute_density(ft: FourierTransform, c, weight: float, occ, omega: float, rho):
    """SlaterDet::compute_density(ft, weight, rho): rho += weight*occ_n/omega |psi_n|^2 (SlaterDet.cc:839-932)"""
    nst, ldc = _block_dims(c)
    fac = np.ascontiguousarray((weight / omega) * np.asarray(occ, dtype=np.float64))
    assert fac.shape[0] == nst
    capi._check(ft._L.qb200_compute_density(ft._h, ldc, nst, capi.ptr(c), capi.ptr(fac), capi.ptr(rho)), "qb200_compute_density")
    return rho


class ChargeDensity:
    """ChargeDensity (ChargeDensity.cc:66-151, 276-556), norm-conserving, one spin / one k-point per object:
    update_density() = zero rho; SlaterDet::compute_density; sum over the ranks that share the states (:309);
    nelectrons = sum(rho)*omega/N (:520-528); rhog = vft->forward(omega*rho) (:550).
    `ft` transforms the wavefunction basis, `vbasis` is the density basis (k = 0, 4*ecut) on the same grid."""

    def __init__(self, ft: FourierTransform, vbasis, omega: float, device: int = 0, stream=None):
        self.ft = ft
        self.vft = FourierTransform(vbasis, ft.np0(), ft.np1(), ft.np2(), device=device, stream=stream)
        self.omega = float(omega)
        self.nelectrons = 0.0

    def update_density(self, c, occ, rhor, rhog, weight: float = 1.0, group=None, ultrasoft=None) -> float:
        """rhor (N doubles) and rhog (vbasis ngw complex) are outputs; returns nelectrons (total_electronic_charge).
        ultrasoft: a NonLocalPotential holding the betag / Q_nm(G) tables -- adds the augmentation charges (:312-465; here before
        the sum over ranks: the term is a sum over states, the reference sums its coefficient matrix over the ranks instead)"""
        from . import parallel as _par
        if hasattr(rhor, "zero_"):
            rhor.zero_()
        else:
            rhor[...] = 0.0
        compute_density(self.ft, c, weight, occ, self.omega, rhor)
        self.uscharge = ultrasoft.us_augment_density(self.vft, c, weight, occ, self.omega, rhor) if ultrasoft is not None else 0.0
        _par.allreduce_density(rhor, group)                        # wfcontext->dsum('r', ...) (:309); host or device array
        nel = C.c_double(0.0)
        capi._check(self.ft._L.qb200_density_finish(self.vft._h, capi.ptr(rhor), self.omega, capi.ptr(rhog), C.byref(nel)),
                    "qb200_density_finish")
        self.nelectrons = nel.value
        return nel.value


XC_LDA, XC_PBE = 0, 1


def update_vhxc(vft: FourierTransform, xc: int, rhor, rhog, gx, g2i, vion_local_g, rhopst, omega: float, v_r, rhogt=None):
    """EnergyFunctional::update_vhxc (EnergyFunctional.cc:353-975) with XCPotential::update (XCPotential.cc:104-460) for one
    spin on the density-basis transform `vft`: v_r (output) = v_xc + FT^-1[vion_local_g + 4 pi (rhog/omega + rhopst) g2i];
    returns (exc, eps, ehart).  xc: XC_LDA or XC_PBE (gx = vbasis.gx_ptr(0), only read for PBE)."""
    en = np.zeros(3)
    capi._check(vft._L.qb200_update_vhxc(vft._h, int(xc), capi.ptr(rhor), capi.ptr(rhog), capi.ptr(gx), capi.ptr(g2i),
                                         capi.ptr(vion_local_g), capi.ptr(rhopst), float(omega), capi.ptr(v_r), capi.ptr(rhogt),
                                         capi.ptr(en)), "qb200_update_vhxc")
    return float(en[0]), float(en[1]), float(en[2])


def compute_current(ft: FourierTransform, c, weight: float, occ, omega: float, kpgx, cur):
    """the per-(spin, k-point) body of CurrentDensity::update_current (CurrentDensity.cc:64-88):
    cur[idir] += -Im sum_n weight*occ_n/omega conj(psi_n) FT^-1[i kpgx_idir c_n]; cur: (3, N) doubles, accumulated."""
    nst, ldc = _block_dims(c)
    fac = np.ascontiguousarray((weight / omega) * np.asarray(occ, dtype=np.float64))
    assert fac.shape[0] == nst
    capi._check(ft._L.qb200_compute_current(ft._h, ldc, nst, capi.ptr(c), capi.ptr(fac), capi.ptr(kpgx), capi.ptr(cur)),
                "qb200_compute_current")
    return cur


class CurrentDensity:
    """CurrentDensity::update_current (CurrentDensity.cc:40-102), one spin / one k-point, without the vector-potential
    term (energy_functional.vp, :96): current[idir][r] and total_current[idir] = volume_element * sum_r current."""

    def __init__(self, ft: FourierTransform, omega: float):
        self.ft, self.omega = ft, float(omega)
        self.total_current = [0.0, 0.0, 0.0]

    def update_current(self, c, occ, kpgx, cur, weight: float = 1.0, group=None):
        from . import parallel as _par
        if hasattr(cur, "zero_"):
            cur.zero_()
        else:
            cur[...] = 0.0
        compute_current(self.ft, c, weight, occ, self.omega, kpgx, cur)
        _par.allreduce_density(cur, group)                         # wfcontext()->dsum('r', ...) (:90); host or device array
        dv = self.omega / self.ft.np012()
        tot = cur.sum(dim=1) if hasattr(cur, "is_cuda") else cur.sum(axis=1)
        self.total_current = [float(dv * t) for t in tot]
        return self.total_current


class NonLocalPotential:
    """NonLocalPotential(atoms, ctxt, basis, ...) norm-conserving branch (NonLocalPotential.cc:76-258, 1909-2171).
    `species` = list of dict(na, npr, lproj, wt, twnl[npr, ngw], tau[na, 3]) -- the reference's init/update_twnl outputs."""

    def __init__(self, basis, species, device: int = 0, stream=None, use_lattice: bool = True):
        L = capi.load()
        g = basis.get if isinstance(basis, dict) else (lambda k, d=None: getattr(basis, k, d))
        kpgx = np.ascontiguousarray(g("kpgx"), dtype=np.float64)
        h = C.c_void_p()
        capi._check(L.qb200_nl_create(C.byref(h), device, int(g("ngw")), int(bool(g("is_real"))), float(g("omega")),
                                      capi.ptr(kpgx)), "qb200_nl_create")
        self._h, self._L = h, L
        if use_lattice and g("idx") is not None and g("cell") is not None and g("kpoint") is not None:
            # integer (h,k,l) + reciprocal lattice: separable phase tables instead of one sincos per (atom, G)
            from . import basis as _B
            idx = np.ascontiguousarray(g("idx"), dtype=np.int32)
            bvec = np.ascontiguousarray(_B.reciprocal_vectors(g("cell")), dtype=np.float64)
            kp = np.ascontiguousarray(g("kpoint"), dtype=np.float64)
            capi._check(L.qb200_nl_set_lattice(h, idx.ctypes.data_as(C.POINTER(C.c_int)), capi.ptr(bvec), capi.ptr(kp)),
                        "qb200_nl_set_lattice")
        for s in species:
            lproj, lp = capi._iarr(s["lproj"])
            wt = np.ascontiguousarray(s["wt"], dtype=np.float64)
            twnl = None if s.get("twnl") is None else np.ascontiguousarray(s["twnl"], dtype=np.float64)   # None: update_twnl() fills it
            tau = np.ascontiguousarray(s["tau"], dtype=np.float64)
            capi._check(L.qb200_nl_add_species(h, int(s["na"]), int(s["npr"]), lp, capi.ptr(wt), capi.ptr(twnl), capi.ptr(tau)),
                        "qb200_nl_add_species")
        if stream is not None:
            s_ = stream.cuda_stream if hasattr(stream, "cuda_stream") else int(stream)
            capi._check(L.qb200_nl_set_stream(h, s_), "qb200_nl_set_stream")

    def launches(self): return int(self._L.qb200_nl_query(self._h, 9))
    def query(self, what): return int(self._L.qb200_nl_query(self._h, int(what)))

    def set_workspace(self, nbytes: int):
        capi._check(self._L.qb200_nl_set_workspace(self._h, int(nbytes)), "qb200_nl_set_workspace")

    def set_positions(self, isp: int, tau):
        tau = np.ascontiguousarray(tau, dtype=np.float64)
        capi._check(self._L.qb200_nl_set_positions(self._h, isp, capi.ptr(tau)), "qb200_nl_set_positions")

    def last_enl(self, out=None):
        """E_nl of the last energy / hpsi call: into a device tensor (1 double, asynchronous) or returned as a float (synchronous)"""
        if out is not None:
            capi._check(self._L.qb200_nl_last_enl(self._h, capi.ptr(out)), "qb200_nl_last_enl")
            return out
        e = np.zeros(1)
        capi._check(self._L.qb200_nl_last_enl(self._h, capi.ptr(e)), "qb200_nl_last_enl")
        return float(e[0])

    def update_twnl(self, isp: int, mproj, tabproj, gspl, vnlg, vnlg_spl, gcut=None):
        """NonLocalPotential::update_twnl for Kleinman-Bylander species `isp` on the device (NonLocalPotential.cc:261-1522, the twnl
        part): mproj / tabproj = m and radial-table index of every projector, gspl / vnlg / vnlg_spl = knots, values and second
        derivatives of the species' radial splines (Species::projectors_g_), gcut = last knot of the full table"""
        m, mp = capi._iarr(mproj)
        t, tp = capi._iarr(tabproj)
        g = np.ascontiguousarray(gspl, dtype=np.float64)
        y = np.ascontiguousarray(vnlg, dtype=np.float64)
        y2 = np.ascontiguousarray(vnlg_spl, dtype=np.float64)
        assert y.ndim == 2 and y.shape == y2.shape and y.shape[1] == g.shape[0]
        capi._check(self._L.qb200_nl_update_twnl(self._h, int(isp), mp, tp, int(y.shape[0]), int(g.shape[0]), capi.ptr(g),
                                                 float(g[-1] if gcut is None else gcut), capi.ptr(y), capi.ptr(y2)), "qb200_nl_update_twnl")

    def update_twnl_semilocal(self, isp: int, mproj, rproj):
        """NonLocalPotential::update_twnl for semi-local species `isp` (nquad > 0) on the device: mproj / rproj = m and quadrature
        radius of every projector (ipr = iquad + nquad * ilm)"""
        m, mp = capi._iarr(mproj)
        r = np.ascontiguousarray(rproj, dtype=np.float64)
        capi._check(self._L.qb200_nl_update_twnl_semilocal(self._h, int(isp), mp, capi.ptr(r)), "qb200_nl_update_twnl_semilocal")

    def get_twnl(self, isp: int, npr: int, ngw: int):
        out = np.zeros((npr, ngw))
        capi._check(self._L.qb200_nl_get_twnl(self._h, int(isp), capi.ptr(out)), "qb200_nl_get_twnl")
        return out

    def energy(self, c, occ, compute_hpsi: bool, cp=None) -> float:
        """energy(sd, compute_hpsi, dsd, ...): returns enl; cp += V_nl psi when compute_hpsi."""
        nst, ldc = _block_dims(c)
        occ = np.ascontiguousarray(occ, dtype=np.float64)
        enl = C.c_double(0.0)
        capi._check(self._L.qb200_nl_energy(self._h, ldc, nst, capi.ptr(c), capi.ptr(occ), int(compute_hpsi), capi.ptr(cp),
                                            C.byref(enl)), "qb200_nl_energy")
        return enl.value

    # ---- ultrasoft beta.psi path: the object was built from the reference's betag tables
    #      (species[i]["twnl"] = beta_b(|k+G|) Y_lm(k+G), "lproj" = l per channel, "wt" unused)
    def nproj(self): return int(self._L.qb200_nl_query(self._h, 13))

    def betapsi(self, c, out):
        """SlaterDet::calc_betapsi (SlaterDet.cc:2130-2263): out[n, p] = <beta_p|psi_n>, p = species, atom, channel"""
        nst, ldc = _block_dims(c)
        capi._check(self._L.qb200_nl_betapsi(self._h, ldc, nst, capi.ptr(c), capi.ptr(out)), "qb200_nl_betapsi")
        return out

    def add_beta(self, f, cp):
        """cp_n(G) += sum_p beta_p(G) f[n, p] (the gemm of SlaterDet::calc_spsi, SlaterDet.cc:2565)"""
        nst, ldc = _block_dims(cp)
        capi._check(self._L.qb200_nl_add_beta(self._h, ldc, nst, capi.ptr(f), capi.ptr(cp)), "qb200_nl_add_beta")
        return cp

    def us_set_tables(self, vkpgx, species):
        """the augmentation tables of the ultrasoft energy branch / augmentation charges: vkpgx = kpgx of the density basis
        (3 x ngv), species = per species a dict with lm1, lm2 (Species::qnm_lm1/2), dzero (Species::dzero) and qnmg (nq x ngv
        complex, Species::calc_qnmg on the density basis; NonLocalPotential.cc:2719, ChargeDensity.cc:793)"""
        vk = np.ascontiguousarray(vkpgx, dtype=np.float64)
        ngv = vk.shape[1]
        capi._check(self._L.qb200_nl_us_set_density_basis(self._h, ngv, capi.ptr(vk)), "qb200_nl_us_set_density_basis")
        for i, s in enumerate(species):
            lm1 = np.ascontiguousarray(s["lm1"], dtype=np.int32)
            lm2 = np.ascontiguousarray(s["lm2"], dtype=np.int32)
            dz = np.ascontiguousarray(s["dzero"], dtype=np.float64)
            q = np.ascontiguousarray(s["qnmg"], dtype=np.complex128)
            assert q.shape == (lm1.shape[0], ngv)
            capi._check(self._L.qb200_nl_us_set_species(self._h, i, int(lm1.shape[0]), lm1.ctypes.data_as(capi.C.POINTER(capi.C.c_int)),
                                                        lm2.ctypes.data_as(capi.C.POINTER(capi.C.c_int)), capi.ptr(dz), capi.ptr(q)),
                        "qb200_nl_us_set_species")

    def us_energy(self, c, occ, veff=None, compute_hpsi=False, cp=None) -> float:
        """ultrasoft branch of NonLocalPotential::energy (NonLocalPotential.cc:1554-1752): returns E_nl; compute_hpsi: cp +=
        beta (D^I <beta|psi>) / omega with D^I = D^0 + sum_G Re(conj(sf_I Q) veff)"""
        nst, ldc = _block_dims(c)
        o = np.ascontiguousarray(occ, dtype=np.float64)
        e = np.zeros(1)
        capi._check(self._L.qb200_nl_us_energy(self._h, ldc, nst, capi.ptr(c), capi.ptr(o), capi.ptr(veff), int(bool(compute_hpsi)),
                                               capi.ptr(cp), capi.ptr(e)), "qb200_nl_us_energy")
        return float(e[0])

    def us_augment_density(self, vft, c, weight: float, occ, omega: float, rho) -> float:
        """augmentation charges of ChargeDensity::update_density (ChargeDensity.cc:312-465): rho += Re FT^-1[rhogus] on the
        density-basis transform `vft`; returns the integrated augmentation charge the reference prints"""
        nst, ldc = _block_dims(c)
        fac = np.ascontiguousarray((weight / omega) * np.asarray(occ, dtype=np.float64))
        q = np.zeros(1)
        capi._check(self._L.qb200_nl_us_augment_density(self._h, vft._h, ldc, nst, capi.ptr(c), capi.ptr(fac), capi.ptr(rho), capi.ptr(q)),
                    "qb200_nl_us_augment_density")
        return float(q[0])

    def spsi(self, c, qmats, out, betapsi=None):
        """SlaterDet::calc_spsi (SlaterDet.cc:2426-2570): out = c + sum beta (q <beta|psi>) / omega; qmats = list of the species'
        dense symmetric npr x npr coupling matrices"""
        nst, ldc = _block_dims(c)
        q = np.ascontiguousarray(np.concatenate([np.asarray(m, dtype=np.float64).ravel() for m in qmats]))
        capi._check(self._L.qb200_nl_spsi(self._h, ldc, nst, capi.ptr(c), capi.ptr(q), capi.ptr(out), capi.ptr(betapsi)), "qb200_nl_spsi")
        return out

    def close(self):
        if getattr(self, "_h", None):
            self._L.qb200_nl_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def hpsi(ft: FourierTransform, nlp, c, occ, v, kpg2, out, want_enl: bool = True):
    """The H psi block of EnergyFunctional::energy(compute_hpsi=true) (EnergyFunctional.cc:1142-1153,1500,1675-1695):
    out = V_nl c + 0.5|k+G|^2 c + FT[v FT^-1 c]; returns enl.  want_enl=False: no host synchronisation, E_nl stays on the
    device (NonLocalPotential.last_enl) and None is returned."""
    nst, ldc = _block_dims(c)
    occ = np.ascontiguousarray(occ, dtype=np.float64)
    enl = C.c_double(0.0)
    capi._check(ft._L.qb200_hpsi(ft._h, nlp._h if nlp is not None else None, ldc, nst, capi.ptr(c), capi.ptr(occ), capi.ptr(v),
                                 capi.ptr(kpg2), capi.ptr(out), C.byref(enl) if want_enl else None), "qb200_hpsi")
    return enl.value if want_enl else None


def ekin_sums(ft: FourierTransform, c, occ, is_real: bool, kpg2, kpgx=None, fstress=None, dfstress=None, want_psi2sum=False):
    """the kinetic-energy section of EnergyFunctional::energy (EnergyFunctional.cc:1155-1296) for one (spin, k-point):
    returns tsum[14] (tsum[0] = ekin partial before the k-point weight, tsum[7] = econf, the rest the stress sums) and,
    optionally, psi2sum[ngw].  fac = 1 for a real basis (G and -G), 0.5 otherwise (:1184)."""
    nst, ldc = _block_dims(c)
    w = np.ascontiguousarray((1.0 if is_real else 0.5) * np.asarray(occ, dtype=np.float64))
    assert w.shape[0] == nst
    tsum = np.zeros(14)
    p2 = np.zeros(ft.ngw()) if want_psi2sum else None
    capi._check(ft._L.qb200_ekin_sums(ft._h, ldc, nst, capi.ptr(c), capi.ptr(w), capi.ptr(kpg2), capi.ptr(kpgx), capi.ptr(fstress),
                                      capi.ptr(dfstress), capi.ptr(p2), capi.ptr(tsum)), "qb200_ekin_sums")
    return (tsum, p2) if want_psi2sum else tsum


def exponential(ft: FourierTransform, nlp, c, occ, v, kpg2, dt1: float, dt2: float = 0.0, c2=None, order: int = 4):
    """ExponentialWavefunctionStepper::exponential(num_exp, dt1, dt2) with a frozen Hamiltonian
    (ExponentialWavefunctionStepper.cc:51-149): c <- sum_N (-i dt1 H)^N/N! c in place; c2 (optional) <- the dt2 series."""
    nst, ldc = _block_dims(c)
    occ = np.ascontiguousarray(occ, dtype=np.float64)
    capi._check(ft._L.qb200_exponential(ft._h, nlp._h if nlp is not None else None, ldc, nst, capi.ptr(c), capi.ptr(occ), capi.ptr(v),
                                        capi.ptr(kpg2), int(order), float(dt1), float(dt2), capi.ptr(c2)), "qb200_exponential")
    return c


class SubspaceLA:
    """The dense linear algebra the ground-state steppers run between two H psi evaluations:
    residual() = the descent direction of PSDAWavefunctionStepper::update (PSDAWavefunctionStepper.cc:65-84, 264-277),
    gram() = SlaterDet::gram (SlaterDet.cc:1043-1143).  Blocks are (nst, ldc) complex arrays/tensors as elsewhere."""

    def __init__(self, basis, device: int = 0, stream=None):
        L = capi.load()
        g = basis.get if isinstance(basis, dict) else (lambda k: getattr(basis, k))
        h = C.c_void_p()
        capi._check(L.qb200_la_create(C.byref(h), device, int(g("ngw")), int(bool(g("is_real")))), "qb200_la_create")
        self._h, self._L = h, L
        self.is_real = bool(g("is_real"))
        if stream is not None:
            s_ = stream.cuda_stream if hasattr(stream, "cuda_stream") else int(stream)
            capi._check(L.qb200_la_set_stream(h, s_), "qb200_la_set_stream")

    def launches(self): return int(self._L.qb200_la_query(self._h, 9))
    def query(self, what): return int(self._L.qb200_la_query(self._h, int(what)))

    def set_workspace(self, nbytes: int):
        capi._check(self._L.qb200_la_set_workspace(self._h, int(nbytes)), "qb200_la_set_workspace")

    def residual(self, c, hc, a=None):
        """hc <- hc - c (c^H hc); c holds ALL states (nall, ldc), hc this rank's (nst, ldc) columns of H psi;
        a (optional output): (nst, nall), row n = column n of the reference's matrix a."""
        nall, ldc = _block_dims(c)
        nst, ldh = _block_dims(hc)
        assert ldh == ldc
        capi._check(self._L.qb200_residual(self._h, ldc, nall, capi.ptr(c), nst, capi.ptr(hc), capi.ptr(a)), "qb200_residual")
        return hc

    def diag(self, c, hc, eigvec: bool = True):
        """Wavefunction::diag (Wavefunction.cc:1510-1715): eigenvalues (ascending) of c^H (H c); eigvec: c <- c z in place.
        Returns (w, sweeps)."""
        nst, ldc = _block_dims(c)
        w = np.zeros(nst)
        sw = C.c_int(0)
        capi._check(self._L.qb200_diag(self._h, ldc, nst, capi.ptr(c), capi.ptr(hc), int(bool(eigvec)), capi.ptr(w), C.byref(sw)), "qb200_diag")
        return w, sw.value

    def psda_update(self, c, dc, c_last, dc_last, occ, precdiag, extrapolate: bool, comm=None) -> float:
        """the rest of PSDAWavefunctionStepper::update (PSDAWavefunctionStepper.cc:93-225, 281-395) on device-resident
        blocks: dc <- -K dc; Anderson extrapolation with theta from the occupation-weighted dot products (summed over the
        ranks of `comm`, a parallel.Communicator); c, c_last, dc_last updated in place.  Returns theta before clipping."""
        nst, ldc = _block_dims(c)
        occ = np.ascontiguousarray(occ, dtype=np.float64)
        th = C.c_double(0.0)
        capi._check(self._L.qb200_psda_update(self._h, comm._h if comm is not None else None, ldc, nst, capi.ptr(c), capi.ptr(dc),
                                              capi.ptr(c_last), capi.ptr(dc_last), capi.ptr(occ), capi.ptr(precdiag),
                                              int(bool(extrapolate)), C.byref(th)), "qb200_psda_update")
        return th.value

    def gram(self, c):
        """c <- c L^-H with c^H c = L L^H, in place"""
        nst, ldc = _block_dims(c)
        info = C.c_int(0)
        capi._check(self._L.qb200_gram(self._h, ldc, nst, capi.ptr(c), C.byref(info)), "qb200_gram")
        return c

    def gram_overlap(self, c_all, first: int, nst: int, S):
        """band-sharded gram, step 1: S (nall x nall complex, device) = 0 but this rank's columns c_all^H c_all[first:first+nst]"""
        nall, ldc = _block_dims(c_all)
        capi._check(self._L.qb200_gram_overlap(self._h, ldc, nall, capi.ptr(c_all), int(first), int(nst), capi.ptr(S)), "qb200_gram_overlap")
        return S

    def gram_apply(self, c_all, S, first: int, nst: int, c_local):
        """band-sharded gram, step 3: c_local <- c_all (L^-H)[:, first:first+nst] with the summed overlap S = L L^H"""
        nall, ldc = _block_dims(c_all)
        info = C.c_int(0)
        capi._check(self._L.qb200_gram_apply(self._h, ldc, nall, capi.ptr(c_all), capi.ptr(S), int(first), int(nst), capi.ptr(c_local),
                                             C.byref(info)), "qb200_gram_apply")
        return c_local

    def gram_sharded(self, comm, c_all, first: int, nst: int, c_local):
        """SlaterDet::gram over band-sharded states: overlap columns -> sum over the ranks of `comm` -> replicated Cholesky -> this
        rank's orthonormalised columns (SlaterDet.cc:1043-1143 with the states distributed over process columns)"""
        nall, ldc = _block_dims(c_all)
        info = C.c_int(0)
        capi._check(self._L.qb200_gram_sharded(self._h, comm._h if comm is not None else None, ldc, nall, capi.ptr(c_all), int(first),
                                               int(nst), capi.ptr(c_local), C.byref(info)), "qb200_gram_sharded")
        return c_local

    def close(self):
        if getattr(self, "_h", None):
            self._L.qb200_la_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
