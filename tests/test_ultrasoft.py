"""the ultrasoft beta.psi path -- SlaterDet::calc_betapsi and SlaterDet::calc_spsi
(/root/reference/src/qball/SlaterDet.cc:2130-2263, 2426-2570).  Fixtures tests/golden/us/*.npz were produced by the reference's
own SlaterDet::init_usfns on the ultrasoft carbon potential of its test suite (tests/golden/make_golden_us.py); the betag
tables in them are the reference's (an input of the C ABI, like twnl)."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(HERE, ".."), os.path.join(HERE, "..", "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)
import port as P  # noqa: E402
import refdrive as R  # noqa: E402
from util import TOL, relerr  # noqa: E402

US = os.path.join(HERE, "golden", "us")
NAMES = sorted(f[:-4] for f in os.listdir(US) if f.endswith(".npz"))


def load(name):
    z = np.load(os.path.join(US, name + ".npz"))
    g = {k: z[k] for k in z.files}
    b = P.make_basis(tuple(g["cell"]), float(g["ecut"]), tuple(g["kpoint"]), bool(int(g["force_complex"])))
    assert b["ngw"] == int(g["ngw"]) and not b["is_real"]
    sp = [dict(na=int(g[f"sp{i}_na"]), npr=int(g[f"sp{i}_npr"]), lproj=g[f"sp{i}_lproj"], twnl=g[f"sp{i}_twnl"], tau=g[f"sp{i}_tau"],
               lm1=g[f"sp{i}_lm1"], lm2=g[f"sp{i}_lm2"], qaug=g[f"sp{i}_qaug"], betapsi=g[f"sp{i}_betapsi"],
               wt=np.zeros(int(g[f"sp{i}_npr"]))) for i in range(int(g["nsp"]))]
    c = R.synth_coefficients(b["kpg2"], float(g["ecut"]), int(g["nst"]), int(g["mloc"]), False, int(g["seed"]))
    return g, b, sp, c


@pytest.mark.parametrize("name", NAMES)
def test_oracle_ultrasoft_betapsi_spsi_vs_reference_fixture(name):
    """the numpy restatement (oracle/port.py us_*) against what the reference's calc_betapsi / calc_spsi produced"""
    g, b, sp, c = load(name)
    spsi, bp = P.us_spsi(b, c, sp)
    want_bp = np.concatenate([s["betapsi"] for s in sp], axis=1)
    assert relerr(bp, want_bp) < 1e-13
    assert relerr(spsi, g["spsi"]) < 1e-13
    for s in sp:                                   # the coupling the reference applies is symmetric
        q = P.us_qmatrix(s)
        assert np.array_equal(q, q.T)


def test_oracle_ultrasoft_vs_live_reference():
    """where the compiled reference is present: a fresh run of ref_driver `us` against the restatement"""
    if not R.have_ref() or not os.path.isdir("/root/reference"):
        pytest.skip("compiled reference not present")
    case = R.Case(cell=(10, 0, 0, 0, 11, 0, 0, 0, 9), ecut=9.0, kpoint=(0.0, 0.5, 0.0), nst=2,
                  species=[("carbon", "/root/reference/testsuite/pseudopotentials/04_ultrasoft_carbon/carbon.xml")],
                  atoms=[("C1", "carbon", 1.0, -2.0, 0.5)])
    r = R.run_reference_us(case, seed=11)
    spsi, bp = P.us_spsi(r, r["c"], r["species"])
    assert relerr(bp, r["species"][0]["betapsi"]) < 1e-13 and relerr(spsi, r["spsi"]) < 1e-13


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("host", [False, True])
def test_cuda_ultrasoft_betapsi_spsi_vs_reference_fixture(name, host):
    """qb200_nl_betapsi / qb200_nl_spsi / qb200_nl_add_beta through the C ABI against the reference's arrays, device and host
    pointers, whole sphere and a workspace that forces several plane-wave chunks"""
    import torch
    from qball_b200 import host as H
    g, b, sp, c = load(name)
    nst = c.shape[0]
    want_bp = np.concatenate([s["betapsi"] for s in sp], axis=1)
    qm = [P.us_qmatrix(s) for s in sp]
    wrap = (lambda a: np.ascontiguousarray(a)) if host else (lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda())
    back = (lambda a: a) if host else (lambda t: t.cpu().numpy())
    for ws in (None, 1 << 20):
        nlp = H.NonLocalPotential(b, sp)
        if ws:
            nlp.set_workspace(ws)
        M = nlp.nproj()
        assert M == want_bp.shape[1]
        bp = wrap(np.zeros((nst, M), dtype=np.complex128))
        nlp.betapsi(wrap(c), bp)
        assert relerr(back(bp), want_bp) < TOL
        out = wrap(np.zeros_like(c))
        bp2 = wrap(np.zeros((nst, M), dtype=np.complex128))
        nlp.spsi(wrap(c), qm, out, bp2)
        assert relerr(back(out)[:, :b["ngw"]], g["spsi"][:, :b["ngw"]]) < TOL
        assert relerr(back(bp2), want_bp) < TOL
        # add_beta alone: cp += anl f against the restatement, on top of a non-zero cp
        rng = np.random.default_rng(5)
        f = rng.standard_normal((nst, M)) + 1j * rng.standard_normal((nst, M))
        cp0 = 0.5 * c
        want = P.us_add_beta(b, f, sp, cp0.copy())
        cp = wrap(cp0.copy())
        nlp.add_beta(wrap(f), cp)
        assert relerr(back(cp)[:, :b["ngw"]], want[:, :b["ngw"]]) < TOL
        # the norm-conserving entry point of the same object still works afterwards (W is regenerated in its own layout)
        e1 = nlp.energy(wrap(c), np.ones(nst), False)
        e2 = nlp.energy(wrap(c), np.ones(nst), False)
        assert e1 == e2
        nlp.close()


@pytest.mark.gpu
def test_cuda_ultrasoft_needs_complex_basis():
    from qball_b200 import capi, host as H
    b = P.make_basis((9, 0, 0, 0, 9, 0, 0, 0, 9), 5.0, (0, 0, 0), False)
    assert b["is_real"]
    sp = [dict(na=1, npr=1, lproj=np.zeros(1, dtype=np.int32), wt=np.zeros(1), twnl=np.ones((1, b["ngw"])), tau=np.zeros((1, 3)))]
    nlp = H.NonLocalPotential(b, sp)
    c = np.zeros((1, b["ngw"]), dtype=np.complex128)
    with pytest.raises(capi.QB200Error):
        nlp.betapsi(c, np.zeros((1, 1), dtype=np.complex128))


# ------------------------------------------------------------------------------------------------ the rest of row f4
# the ultrasoft branch of NonLocalPotential::energy (NonLocalPotential.cc:1554-1752) and the augmentation charges of
# ChargeDensity::update_density (ChargeDensity.cc:312-465); fixtures tests/golden/usx/*.npz from the reference's own classes on a
# Sample (tests/golden/make_golden_usx.py)
USX = os.path.join(HERE, "golden", "usx")
XNAMES = sorted(f[:-4] for f in os.listdir(USX) if f.endswith(".npz"))


def loadx(name):
    z = np.load(os.path.join(USX, name + ".npz"))
    g = {k: z[k] for k in z.files}
    cell, ecut = tuple(g["cell"]), float(g["ecut"])
    b = P.make_basis(cell, ecut, tuple(g["kpoint"]), True)
    vb = P.make_basis(cell, 4.0 * ecut, (0, 0, 0), True)        # ChargeDensity.cc:75-80: complex at k = 0 for ultrasoft runs
    assert b["ngw"] == int(g["ngw"]) and vb["ngw"] == int(g["ngv"]) and P.density_grid(cell, ecut) == tuple(g["vgrid"])
    sp = [dict(na=int(g[f"sp{i}_na"]), npr=int(g[f"sp{i}_npr"]), nq=int(g[f"sp{i}_nq"]), lproj=g[f"sp{i}_lproj"], twnl=g[f"sp{i}_twnl"],
               tau=g[f"sp{i}_tau"], lm1=g[f"sp{i}_lm1"], lm2=g[f"sp{i}_lm2"], dzero=g[f"sp{i}_dzero"], qnmg=g[f"sp{i}_qnmg"],
               betapsi=g[f"sp{i}_betapsi"], wt=np.zeros(int(g[f"sp{i}_npr"]))) for i in range(int(g["nsp"]))]
    c = R.synth_coefficients(b["kpg2"], ecut, int(g["nst"]), int(g["mloc"]), False, int(g["seed"]))
    return g, b, vb, sp, c


@pytest.mark.parametrize("name", XNAMES)
def test_oracle_ultrasoft_energy_and_augmentation_vs_reference_fixture(name):
    """the numpy restatement (oracle/port.py us_energy / us_augment_density) against what the reference's
    NonLocalPotential::energy and ChargeDensity::update_density produced"""
    g, b, vb, sp, c = loadx(name)
    ngw = b["ngw"]
    enl, hp, bp = P.us_energy(b, c, g["occ"], sp, vb["kpgx"], g["veff"])
    assert relerr(bp, np.concatenate([s["betapsi"] for s in sp], axis=1)) < 1e-13
    assert abs(enl - float(g["enl"])) < 1e-13 * abs(float(g["enl"]))
    assert relerr(hp[:, :ngw], g["hnl"][:, :ngw]) < 1e-13
    vft = P.FT(vb, *(int(x) for x in g["vgrid"]))
    rho, usq = P.us_augment_density(b, vft, c, g["occ"] / b["omega"], sp, vb["kpgx"], g["rho_nc"].copy())
    assert relerr(rho, g["rho"]) < 1e-13
    assert np.abs(g["rho"] - g["rho_nc"]).max() > 0.1 * np.abs(g["rho"]).max()       # the augmentation part is not small here
    assert abs(rho.sum() * b["omega"] / rho.size - float(g["nelectrons"])) < 1e-11


def test_oracle_ultrasoft_energy_vs_live_reference():
    if not R.have_ref() or not os.path.isdir("/root/reference"):
        pytest.skip("compiled reference not present")
    case = R.Case(cell=(9, 0, 0, 0, 8, 0, 0.4, 0, 8.5), ecut=4.0, kpoint=(0.0, 0.5, 0.0), nst=2,
                  species=[("carbon", "/root/reference/testsuite/pseudopotentials/04_ultrasoft_carbon/carbon.xml")],
                  atoms=[("C1", "carbon", 1.0, -2.0, 0.5)])
    r = R.run_reference_usx(case, seed=11, occ=[2.0, 0.3])
    enl, hp, _ = P.us_energy(r, r["c"], r["occ"], r["species"], r["vkpgx"], r["veff"])
    assert abs(enl - r["enl"]) < 1e-13 * abs(r["enl"]) and relerr(hp[:, :r["ngw"]], r["hnl"][:, :r["ngw"]]) < 1e-13
    vb = P.make_basis(case.cell, 4.0 * case.ecut, (0, 0, 0), True)
    rho, _ = P.us_augment_density(r, P.FT(vb, *r["vgrid"]), r["c"], r["occ"] / r["omega"], r["species"], r["vkpgx"], r["rho_nc"].copy())
    assert relerr(rho, r["rho"]) < 1e-13


@pytest.mark.gpu
@pytest.mark.parametrize("name", XNAMES)
@pytest.mark.parametrize("host", [False, True])
def test_cuda_ultrasoft_energy_and_augmentation_vs_reference_fixture(name, host):
    """qb200_nl_us_energy / qb200_nl_us_augment_density through the C ABI against the reference's arrays (E_nl, H psi with the
    atom-dependent D^I built from veff(G) and Q_nm(G), rho with the augmentation charges), device and host pointers, whole
    sphere and a workspace that forces several plane-wave chunks"""
    import torch
    from qball_b200 import host as H
    g, b, vb, sp, c = loadx(name)
    ngw, nst = b["ngw"], c.shape[0]
    wrap = (lambda a: np.ascontiguousarray(a)) if host else (lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda())
    back = (lambda a: a) if host else (lambda t: t.cpu().numpy())
    vgrid = tuple(int(x) for x in g["vgrid"])
    for ws in (None, 1 << 20):
        nlp = H.NonLocalPotential(b, sp)
        if ws:
            nlp.set_workspace(ws)
        nlp.us_set_tables(vb["kpgx"], sp)
        e0 = nlp.us_energy(wrap(c), g["occ"])                               # energy only
        assert abs(e0 - float(g["enl"])) < TOL * abs(float(g["enl"]))
        cp0 = 0.25 * c
        cp = wrap(cp0.copy())
        e1 = nlp.us_energy(wrap(c), g["occ"], wrap(g["veff"]), True, cp)    # H psi on top of a non-zero cp
        assert e1 == e0
        assert relerr(back(cp)[:, :ngw] - cp0[:, :ngw], g["hnl"][:, :ngw]) < TOL
        assert np.all(back(cp)[:, ngw:] == cp0[:, ngw:])
        # density: SlaterDet::compute_density on the wavefunction basis, then the augmentation charges on the density basis
        ft = H.FourierTransform(b, *vgrid)
        vft = H.FourierTransform(vb, *vgrid)
        rho = wrap(np.zeros(vgrid[0] * vgrid[1] * vgrid[2]))
        H.compute_density(ft, wrap(c), 1.0, g["occ"], b["omega"], rho)
        assert relerr(back(rho), g["rho_nc"]) < TOL
        usq = nlp.us_augment_density(vft, wrap(c), 1.0, g["occ"], b["omega"], rho)
        assert relerr(back(rho), g["rho"]) < TOL
        want_q = (g["rho"] - g["rho_nc"]).sum() * b["omega"] / g["rho"].size
        assert abs(usq - want_q) < 1e-10 * max(1.0, abs(want_q))
        assert abs(back(rho).sum() * b["omega"] / g["rho"].size - float(g["nelectrons"])) < 1e-10
        # the same through the ChargeDensity mirror: rho, rho(G) and the electron count the reference reports
        cdm = H.ChargeDensity(ft, vb, b["omega"])
        rho2, rhog = wrap(np.zeros_like(g["rho"])), wrap(np.zeros(vb["ngw"], dtype=np.complex128))
        nel = cdm.update_density(wrap(c), g["occ"], rho2, rhog, ultrasoft=nlp)
        assert relerr(back(rho2), g["rho"]) < TOL and abs(nel - float(g["nelectrons"])) < 1e-10 and abs(cdm.uscharge - want_q) < 1e-10
        assert abs(back(rhog)[np.argmin(vb["kpg2"])].real - float(g["nelectrons"])) < 1e-9     # rho(G = 0) = omega * mean(rho) = nel
        nlp.close()


@pytest.mark.gpu
def test_cuda_ultrasoft_energy_needs_tables():
    from qball_b200 import capi, host as H
    g, b, vb, sp, c = loadx(XNAMES[0])
    nlp = H.NonLocalPotential(b, sp)
    with pytest.raises(capi.QB200Error):
        nlp.us_energy(c, g["occ"])
    nlp.close()
