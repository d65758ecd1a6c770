"""v(r) producers on the density basis: EnergyFunctional::update_vhxc (EnergyFunctional.cc:353-975)
with XCPotential::update (XCPotential.cc:104-460) and the unpolarized LDA / PBE functionals.
CPU: the oracle's functionals against golden vectors produced by the reference's own LDAFunctional / PBEFunctional
(tests/golden/xc/xc_points.npz, bit-exact) and live against the compiled reference; properties of the oracle's update_vhxc.
GPU: qb200_update_vhxc through the C ABI against the oracle.  (The reference-level pin is tests/test_reference_shim.py:
the reference's own SCF runs with update_vhxc forwarded to the device reproduce every printed energy to 1e-8 Ha.)"""
import os

import numpy as np
import pytest

import port as P
import refdrive as R
from util import GOLDEN, TOL, relerr


def _xc_fixture():
    z = np.load(os.path.join(GOLDEN, "xc", "xc_points.npz"))
    return {k: z[k] for k in z.files}


def test_oracle_xc_functionals_match_reference_fixture():
    g = _xc_fixture()
    e, v = P.xc_lda(g["rho"])
    assert np.array_equal(e, g["lda_exc"]) and np.array_equal(v, g["lda_vxc"])
    gm = np.sqrt((g["grad"] ** 2).sum(axis=0))
    e, v1, v2 = P.xc_pbe(g["rho"], gm)
    for got, want in ((e, g["pbe_exc"]), (v1, g["pbe_vxc1"]), (v2, g["pbe_vxc2"])):
        assert np.all(np.abs(got - want) <= 1e-14 * np.abs(want))
    assert np.all(g["pbe_exc"][g["rho"] < 1e-18] == 0.0)                 # PBEFunctional.cc:219-221


@pytest.mark.skipif(not R.have_ref(), reason="oracle/_ref/ref_driver not built (needs /root/reference)")
def test_oracle_xc_functionals_match_reference_live():
    import sys
    sys.path.insert(0, GOLDEN)
    import make_golden_xc as M
    rho, grad = M.points(n=700, seed=21)
    out = M.run_reference(rho, grad)
    e, v = P.xc_lda(rho)
    assert np.array_equal(e, out["lda_exc"]) and np.array_equal(v, out["lda_vxc"])
    e, v1, v2 = P.xc_pbe(rho, np.sqrt((grad ** 2).sum(axis=0)))
    assert np.all(np.abs(e - out["pbe_exc"]) <= 1e-14 * np.abs(out["pbe_exc"]))
    assert np.all(np.abs(v2 - out["pbe_vxc2"]) <= 1e-14 * np.abs(out["pbe_vxc2"]))


def _case(seed=5, ecut=4.0, cell=(9, 0, 0, 0.4, 10, 0, 0, 0, 11)):
    """a smooth positive density on the density grid of a small cell, its rho(G) = FT[omega rho] on the density basis
    (ChargeDensity.cc:523-550), and Gaussian stand-ins for vion_local_g / rhopst"""
    vb = P.make_basis(cell, 4.0 * ecut, (0, 0, 0), False)              # vbasis: k = 0, 4 ecut (ChargeDensity.cc:77-81)
    grid = P.density_grid(cell, ecut)
    b = P.make_basis(cell, ecut, (0, 0, 0), False)
    oft, ovft = P.FT(b, *grid), P.FT(vb, *grid)
    c = R.synth_coefficients(b["kpg2"], ecut, 4, b["ngw"], True, seed=seed)
    rho = oft.compute_density(c, np.array([2.0, 2.0, 1.0, 0.5]) / b["omega"], np.zeros(oft.N)) + 1e-3
    rhog = ovft.forward((vb["omega"] * rho).astype(np.complex128))
    g2 = vb["kpg2"]
    vion = (-3.0 * np.exp(-0.35 * g2) * np.cos(vb["kpgx"][0] * 1.3 + vb["kpgx"][2] * 0.4)).astype(np.complex128) / vb["omega"] * 40.0
    rhopst = (-(7.5 / vb["omega"]) * np.exp(-0.25 * 0.8 ** 2 * g2) * np.exp(-1j * (vb["kpgx"][1] * 0.9))).astype(np.complex128)
    return vb, grid, ovft, rho, rhog, vion, rhopst


def test_oracle_update_vhxc_hartree_and_local_terms():
    """conventions of EnergyFunctional.cc:447-518 on the oracle: with rhopst = 0 the Hartree energy is 1/2 int rho v_H,
    eps = int rho v_ion, and E_xc = int rho exc"""
    vb, grid, ovft, rho, rhog, vion, rhopst = _case()
    N = ovft.N
    dv = vb["omega"] / N
    v0, e0 = P.update_vhxc(ovft, vb, 0, rho, rhog, 0 * vion, 0 * rhopst)
    exc, vxc = P.xc_lda(rho)
    vh = v0 - vxc
    assert abs(e0[0] - float(np.dot(rho, exc)) * dv) <= 1e-12 * abs(e0[0])
    assert abs(e0[2] - 0.5 * float(np.dot(rho, vh)) * dv) <= 1e-10 * abs(e0[2])
    assert abs(e0[1]) == 0.0
    v1, e1 = P.update_vhxc(ovft, vb, 0, rho, rhog, vion, 0 * rhopst)
    assert abs(e1[1] - float(np.dot(rho, v1 - v0)) * dv) <= 1e-10 * abs(e1[1])


@pytest.mark.gpu
@pytest.mark.parametrize("xc,host", [(0, False), (1, False), (1, True), (0, True)])
def test_cuda_update_vhxc_vs_oracle(xc, host):
    import torch
    from qball_b200 import host as H
    vb, grid, ovft, rho, rhog, vion, rhopst = _case()
    v_ref, e_ref = P.update_vhxc(ovft, vb, xc, rho, rhog, vion, rhopst)
    vft = H.FourierTransform(vb, *grid)
    w = (lambda a: np.ascontiguousarray(a)) if host else (lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda())
    back = (lambda a: a) if host else (lambda t: t.cpu().numpy())
    v_r = w(np.zeros(ovft.N))
    rt = w(np.zeros(vb["ngw"], dtype=np.complex128))
    en = H.update_vhxc(vft, xc, w(rho), w(rhog), w(vb["kpgx"]), w(P.g2i_of(vb)), w(vion), w(rhopst), vb["omega"], v_r, rt)
    assert relerr(back(v_r), v_ref) < TOL
    assert np.abs(np.array(en) - e_ref).max() < 1e-10 * np.abs(e_ref).max(), (en, e_ref)
    assert relerr(back(rt), rhog / vb["omega"] + rhopst) < 1e-14
    # a second call with a different density reuses the work space
    en2 = H.update_vhxc(vft, xc, w(1.5 * rho), w(1.5 * rhog), w(vb["kpgx"]), w(P.g2i_of(vb)), w(vion), w(rhopst), vb["omega"], v_r, None)
    v2, e2 = P.update_vhxc(ovft, vb, xc, 1.5 * rho, 1.5 * rhog, vion, rhopst)
    assert relerr(back(v_r), v2) < TOL and np.abs(np.array(en2) - e2).max() < 1e-10 * np.abs(e2).max()
