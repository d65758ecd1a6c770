"""NonLocalPotential::update_twnl (/root/reference/src/qball/NonLocalPotential.cc:261-1522, the twnl part,
Kleinman-Bylander species) -- the projector tables twnl[ipr][ig] = Y_lm(k+G) v(|k+G|) from the species' radial cubic splines
(Species::dvnlg, Species.cc:1492-1505; splintd, src/pseudo/spline.cc:126-156).  The fixtures tests/golden/twnl/*.npz hold the
spline tables of the reference's own Species objects (tests/golden/make_golden_twnl.py); the twnl tables the reference's
update_twnl produced from them are in the main fixtures."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(HERE, ".."), os.path.join(HERE, "..", "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)
import port as P  # noqa: E402
from util import load_golden, regen_inputs, relerr  # noqa: E402

TW = os.path.join(HERE, "golden", "twnl")
NAMES = sorted(f[:-4] for f in os.listdir(TW) if f.endswith(".npz"))


def load(name):
    g = load_golden(name)
    z = np.load(os.path.join(TW, name + ".npz"))
    t = {k: z[k] for k in z.files}
    b = P.make_basis(g["cell"], g["ecut"], tuple(g["kpoint"]), bool(g["force_complex"]))
    assert b["ngw"] == g["ngw"] and int(t["nsp"]) == g["nsp"]
    return g, t, b


@pytest.mark.parametrize("name", NAMES)
def test_oracle_update_twnl_bit_exact_vs_reference_tables(name):
    """the numpy restatement (oracle/port.py update_twnl) against the tables the reference's update_twnl produced: l = 0..3,
    two-channel ONCV, a k-point, a triclinic cell, the MgO216 basis -- identical to the last bit (same arithmetic order)"""
    g, t, b = load(name)
    seen = 0
    for i, s in enumerate(g["species"]):
        if not int(t[f"sp{i}_kb"]):
            continue
        got = P.update_twnl(b["kpgx"], s["lproj"], t[f"sp{i}_mproj"], t[f"sp{i}_tabproj"], t[f"sp{i}_gspl"], t[f"sp{i}_vnlg"],
                            t[f"sp{i}_vnlg_spl"], gcut=float(t[f"sp{i}_gcut"]))
        assert np.array_equal(got, s["twnl"])
        seen += 1
    assert seen >= 1


def test_oracle_update_twnl_zero_beyond_the_last_knot():
    g, t, b = load(NAMES[0])
    s = g["species"][0]
    gc = float(np.sqrt(b["kpg2"].max())) * 0.6                      # a cut-off inside the sphere: v = 0 beyond it (Species.cc:1495)
    got = P.update_twnl(b["kpgx"], s["lproj"], t["sp0_mproj"], t["sp0_tabproj"], t["sp0_gspl"], t["sp0_vnlg"], t["sp0_vnlg_spl"], gcut=gc)
    far = np.sqrt((b["kpgx"] ** 2).sum(axis=0)) > gc
    assert far.any() and np.all(got[:, far] == 0.0) and np.array_equal(got[:, ~far], s["twnl"][:, ~far])


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_cuda_update_twnl_vs_reference_tables(name):
    """qb200_nl_update_twnl fills the species' tables on the device from the radial splines; the tables (read back with
    qb200_nl_get_twnl) match the reference's to 1e-13 of their largest element, and E_nl / V_nl psi computed from them match
    the fixture exactly as with the reference's tables handed over"""
    import torch
    from qball_b200 import host as H
    g, t, b = load(name)
    c, v, occ = regen_inputs(g, b["kpg2"])
    species = [dict(s, twnl=None) if int(t[f"sp{i}_kb"]) else s for i, s in enumerate(g["species"])]
    nlp = H.NonLocalPotential(b, species)
    for i, s in enumerate(g["species"]):
        if not int(t[f"sp{i}_kb"]):
            continue
        assert np.all(nlp.get_twnl(i, s["npr"], g["ngw"]) == 0.0)
        nlp.update_twnl(i, t[f"sp{i}_mproj"], t[f"sp{i}_tabproj"], t[f"sp{i}_gspl"], t[f"sp{i}_vnlg"], t[f"sp{i}_vnlg_spl"],
                        gcut=float(t[f"sp{i}_gcut"]))
        assert relerr(nlp.get_twnl(i, s["npr"], g["ngw"]), s["twnl"]) < 1e-13
    cd = torch.from_numpy(c).cuda()
    cp = torch.zeros_like(cd)
    enl = nlp.energy(cd, occ, True, cp)
    ref = H.NonLocalPotential(b, g["species"])
    cp2 = torch.zeros_like(cd)
    enl2 = ref.energy(cd, occ, True, cp2)
    assert abs(enl - g["enl"]) <= 1e-10 * max(1.0, abs(g["enl"])) and abs(enl - enl2) <= 1e-12 * max(1.0, abs(enl2))
    assert relerr(cp.cpu().numpy(), cp2.cpu().numpy()) < 1e-12
    # a second update (cell change) invalidates every cached anl: same answer again
    for i, s in enumerate(g["species"]):
        if int(t[f"sp{i}_kb"]):
            nlp.update_twnl(i, t[f"sp{i}_mproj"], t[f"sp{i}_tabproj"], t[f"sp{i}_gspl"], t[f"sp{i}_vnlg"], t[f"sp{i}_vnlg_spl"],
                            gcut=float(t[f"sp{i}_gcut"]))
    assert nlp.energy(cd, occ, False) == enl
    nlp.close(); ref.close()


@pytest.mark.gpu
def test_cuda_update_twnl_rejects_bad_descriptions():
    from qball_b200 import capi, host as H
    g, t, b = load(NAMES[0])
    s = g["species"][0]
    nlp = H.NonLocalPotential(b, [dict(s, twnl=None)])
    with pytest.raises(capi.QB200Error):
        nlp.update_twnl(0, t["sp0_mproj"] + 7, t["sp0_tabproj"], t["sp0_gspl"], t["sp0_vnlg"], t["sp0_vnlg_spl"])
    with pytest.raises(capi.QB200Error):
        nlp.update_twnl(0, t["sp0_mproj"], t["sp0_tabproj"], t["sp0_gspl"][::-1].copy(), t["sp0_vnlg"], t["sp0_vnlg_spl"])
    nlp.close()


# ---- semi-local species (nquad > 0): twnl = Y_lm(k+G) 4 pi j_l(|k+G| r_iquad) r_iquad; fixtures tests/golden/twnl_semilocal/*.npz
SL = os.path.join(HERE, "golden", "twnl_semilocal")
SLNAMES = sorted(f[:-4] for f in os.listdir(SL) if f.endswith(".npz"))


def load_sl(name):
    z = np.load(os.path.join(SL, name + ".npz"))
    t = {k: z[k] for k in z.files}
    b = P.make_basis(tuple(t["cell"]), float(t["ecut"]), tuple(t["kpoint"]), False)
    assert b["ngw"] == int(t["ngw"])
    return t, b


@pytest.mark.parametrize("name", SLNAMES)
def test_oracle_update_twnl_semilocal_vs_reference_tables(name):
    """the numpy restatement against the reference's tables for the shipped semi-local potentials (Yb: l = 0, 1, 2; Zr: l = 1, 2;
    8 quadrature radii each): equal to the last bits (sin / cos of numpy and of libm may differ by an ulp)"""
    t, b = load_sl(name)
    got = P.update_twnl_semilocal(b["kpgx"], t["lproj"], t["mproj"], t["rproj"])
    assert relerr(got, t["twnl"]) < 1e-15


def test_oracle_update_twnl_semilocal_l3_is_the_spherical_bessel_function():
    """no shipped semi-local potential has l = 3 projectors: the restatement's j_3 (NonLocalPotential.cc:1305-1313) against scipy"""
    from scipy.special import spherical_jn
    t, b = load_sl(SLNAMES[0])
    r = 1.7
    got = P.update_twnl_semilocal(b["kpgx"], [3] * 7, list(range(7)), [r] * 7)
    q = np.sqrt((b["kpgx"] ** 2).sum(axis=0))
    _, _, ylm = P._ylm_table(b["kpgx"])
    for m in range(7):
        want = ylm[(3, m)] * 4.0 * np.pi * spherical_jn(3, q * r) * r
        big = q * r > 0.5                      # the sin / cos form loses digits at small arguments (as the reference's does)
        assert np.abs(got[m][big] - want[big]).max() < 1e-11 * np.abs(want).max()


@pytest.mark.gpu
@pytest.mark.parametrize("name", SLNAMES)
def test_cuda_update_twnl_semilocal_vs_reference_tables(name):
    """qb200_nl_update_twnl_semilocal fills the tables on the device; read back they match the reference's, and E_nl computed
    from them matches the reference's NonLocalPotential::energy on the fixture's seeded states"""
    import torch
    import refdrive as R
    from qball_b200 import host as H
    t, b = load_sl(name)
    sp = dict(na=int(t["na"]), npr=int(t["npr"]), lproj=t["lproj"], wt=t["wt"], tau=t["tau"].reshape(-1, 3), twnl=None)
    nlp = H.NonLocalPotential(b, [sp])
    nlp.update_twnl_semilocal(0, t["mproj"], t["rproj"])
    # (the sin / cos form of j_l cancels at small q r, in the reference as here: an ulp of sincos is worth ~1e-12 of the table)
    assert relerr(nlp.get_twnl(0, sp["npr"], b["ngw"]), t["twnl"]) < 1e-11
    c = R.synth_coefficients(b["kpg2"], float(t["ecut"]), int(t["nst"]), int(t["mloc"]), b["is_real"], int(t["seed"]))
    occ = R.synth_occ(int(t["nst"]), None)
    cd = torch.from_numpy(c).cuda()
    cp = torch.zeros_like(cd)
    enl = nlp.energy(cd, occ, True, cp)
    assert abs(enl - float(t["enl"])) <= 1e-10 * max(1.0, abs(float(t["enl"])))
    assert abs(float(np.abs(cp.cpu().numpy()).sum()) - float(t["hnl_checksum"])) <= 1e-9 * float(t["hnl_checksum"])
    # l = 3 (no shipped potential): the device against the restatement
    l3 = dict(na=1, npr=7, lproj=np.full(7, 3, dtype=np.int32), wt=np.ones(7), tau=np.zeros((1, 3)), twnl=None)
    n3 = H.NonLocalPotential(b, [l3])
    n3.update_twnl_semilocal(0, np.arange(7), np.full(7, 1.7))
    assert relerr(n3.get_twnl(0, 7, b["ngw"]), P.update_twnl_semilocal(b["kpgx"], [3] * 7, list(range(7)), [1.7] * 7)) < 1e-11
    nlp.close(); n3.close()
