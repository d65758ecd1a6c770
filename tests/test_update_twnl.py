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
