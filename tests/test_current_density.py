"""Current density: CurrentDensity::update_current (CurrentDensity.cc:52-86) around
SlaterDet::compute_density(ft, w, complex rho, sd2) (SlaterDet.cc:935-968).
CPU: the oracle against golden vectors produced by the reference's own pair-density routine (tests/golden/la/*.npz, key
`cur`).  GPU: qb200_compute_current against the same vectors and against the oracle on larger seeded blocks.
Tolerance: 1e-10 of max|conj(psi) grad psi| (the current of REAL wavefunctions is identically zero: compared absolutely)."""
import os

import numpy as np
import pytest

import port as P
import refdrive as R
from util import GOLDEN, TOL, golden_names, load_golden, regen_inputs


def _case(name):
    g = load_golden(name)
    b = P.make_basis(g["cell"], g["ecut"], tuple(g["kpoint"]), bool(g["force_complex"]))
    c, v, occ = regen_inputs(g, b["kpg2"])
    cur = np.load(os.path.join(GOLDEN, "la", name + ".npz"))["cur"]
    return g, b, c, occ, cur


def _scale(cur, c, b, occ):
    # size of the summands: sum_n occ_n * |psi|_max * |grad psi|_max bounds rounding; use max|cur| unless it vanishes
    m = float(np.abs(cur).max())
    return m if m > 1e-8 else 1.0


@pytest.mark.parametrize("name", golden_names("full"))
def test_oracle_current_matches_reference_fixture(name):
    g, b, c, occ, cur_ref = _case(name)
    ft = P.FT(b, g["np0"], g["np1"], g["np2"])
    cur = P.compute_current(ft, c, occ / g["omega"], b["kpgx"])   # prefac = weight/volume with weight 1 (SlaterDet.cc:944)
    assert np.abs(cur - cur_ref).max() < TOL * _scale(cur_ref, c, b, occ)
    if g["is_real"]:
        assert np.abs(cur_ref).max() < 1e-12                  # real wavefunctions carry no current


@pytest.mark.gpu
@pytest.mark.parametrize("device_ptrs", [True, False])
@pytest.mark.parametrize("name", golden_names("full"))
def test_cuda_current_vs_reference_fixture(name, device_ptrs):
    import torch
    from qball_b200 import host as H
    g, b, c, occ, cur_ref = _case(name)
    ft = H.FourierTransform(b, g["np0"], g["np1"], g["np2"])
    N = ft.np012()
    wrap = (lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()) if device_ptrs else (lambda a: np.ascontiguousarray(a).copy())
    back = (lambda t: t.cpu().numpy()) if device_ptrs else (lambda a: a)
    cur = wrap(np.zeros((3, N)))
    H.compute_current(ft, wrap(c), 1.0, occ, g["omega"], wrap(b["kpgx"]), cur)
    assert np.abs(back(cur) - cur_ref).max() < TOL * _scale(cur_ref, c, b, occ)


@pytest.mark.gpu
@pytest.mark.parametrize("kpoint,fc,nst,ldpad", [((0, 0, 0), True, 37, 0), ((0.25, -0.1, 0.5), False, 20, 3)])
def test_cuda_current_vs_oracle_and_total(kpoint, fc, nst, ldpad):
    import torch
    from qball_b200 import host as H
    cell, ecut = (11, 0, 0, 0.5, 10, 0, 0, 0, 12), 6.0
    b = P.make_basis(cell, ecut, kpoint, fc)
    grid = P.density_grid(cell, ecut)
    c = R.synth_coefficients(b["kpg2"], ecut, nst, b["ngw"] + ldpad, b["is_real"], 31)
    occ = R.synth_occ(nst, nst - 3)                             # the last states are empty: fac = 0 -> skipped
    oft = P.FT(b, *grid)
    ref = P.compute_current(oft, c, occ / b["omega"], b["kpgx"])
    ft = H.FourierTransform(b, *grid)
    cd = H.CurrentDensity(ft, b["omega"])
    cur = torch.full((3, oft.N), 7.0, dtype=torch.float64, device="cuda")   # update_current clears it first
    tot = cd.update_current(torch.from_numpy(c).cuda(), occ, torch.from_numpy(b["kpgx"]).cuda(), cur)
    assert np.abs(cur.cpu().numpy() - ref).max() < TOL * np.abs(ref).max()
    # total_electronic_current = volume_element * sum_r current (CurrentDensity.cc:92-94); also = sum_n f_n <psi|k+G|psi>
    want = [float(np.sum(occ[:, None] * np.abs(c[:, :b["ngw"]]) ** 2 * b["kpgx"][d][None, :])) * -1.0 for d in range(3)]
    for d in range(3):
        assert abs(tot[d] - ref[d].sum() * b["omega"] / oft.N) < 1e-10 * max(1.0, abs(want[d]))
        assert abs(tot[d] - want[d]) < 1e-9 * max(1.0, abs(want[d]))


@pytest.mark.gpu
def test_cuda_current_properties_mgo216_full_size():
    """BASELINE-size shape (MgO216: 112^3, ngw 73447): the integral of the current density equals -sum_n f_n <psi_n|k+G|psi_n>
    (exact identity of the definition), j is odd under complex conjugation of the coefficients' mirror image
    c(G) -> conj(c(-G)) (time reversal), which for a state with c(-G) = conj(c(G)) (a real function) means j = 0."""
    import torch
    from qball_b200 import host as H
    cell, ecut, nst = (23.1, 0, 0, 0, 23.1, 0, 0, 0, 23.1), 25.0, 5
    b = P.make_basis(cell, ecut, (0, 0, 0), True)
    grid = P.density_grid(cell, ecut)
    N = grid[0] * grid[1] * grid[2]
    ngw = b["ngw"]
    c = R.synth_coefficients(b["kpg2"], ecut, nst, ngw, False, seed=51)
    # make the last state a REAL function: c(-G) = conj(c(G))
    idx = b["idx"]
    where = {tuple(v): i for i, v in enumerate(idx.tolist())}
    minus = np.array([where[(-h, -k, -l)] for h, k, l in idx.tolist()])
    c[-1] = 0.5 * (c[-1] + np.conj(c[-1][minus]))
    occ = np.array([2.0, 1.0, 0.0, 0.5, 2.0])
    ft = H.FourierTransform(b, *grid)
    cur = torch.zeros((3, N), dtype=torch.float64, device="cuda")
    kd = torch.from_numpy(b["kpgx"]).cuda()
    H.compute_current(ft, torch.from_numpy(c).cuda(), 1.0, occ, b["omega"], kd, cur)
    tot = cur.sum(dim=1).cpu().numpy() * b["omega"] / N
    want = np.array([-float(np.sum(occ[:, None] * np.abs(c) ** 2 * b["kpgx"][d][None, :])) for d in range(3)])
    scale = float(np.sum(occ[:, None] * np.abs(c) ** 2 * np.sqrt(2 * b["kpg2"])[None, :]))
    assert np.abs(tot - want).max() < 1e-10 * scale
    # the real state alone carries no current
    only = np.zeros(nst); only[-1] = 2.0
    cur1 = torch.zeros((3, N), dtype=torch.float64, device="cuda")
    H.compute_current(ft, torch.from_numpy(c).cuda(), 1.0, only, b["omega"], kd, cur1)
    assert float(cur1.abs().max()) < 1e-10 * float(cur.abs().max())


@pytest.mark.skipif(not R.have_ref(), reason="oracle/_ref/ref_driver not built (needs /root/reference)")
@pytest.mark.parametrize("cell,ecut,kpoint,fc,nst", [
    ((10, 0, 0, 0, 10, 0, 0, 0, 10), 6.0, (0, 0, 0), True, 4),
    ((8, 0, 0, 2.0, 9, 0, -1.0, 0.5, 12), 5.0, (0.5, 0.5, 0.5), False, 3),
    ((5.4, 5.4, 0, 0, 5.4, 5.4, 5.4, 0, 5.4), 8.0, (0.125, 0.25, 0.375), False, 2),
])
def test_oracle_current_vs_live_reference(cell, ecut, kpoint, fc, nst):
    """the oracle against the reference's pair-form compute_density run live (shapes the fixtures do not cover:
    triclinic and fcc cells, general k-points), plus the integral identity of the definition"""
    r = R.run_reference(R.Case(cell=cell, ecut=ecut, kpoint=kpoint, force_complex=fc, nst=nst), seed=9, nocc=nst - 1)
    b = P.make_basis(cell, ecut, kpoint, fc)
    ft = P.FT(b, r["np0"], r["np1"], r["np2"])
    cur = P.compute_current(ft, r["c"], r["occ"] / r["omega"], r["kpgx"])
    assert np.abs(cur - r["cur"]).max() < TOL * np.abs(r["cur"]).max()
    ngw = b["ngw"]
    tot = cur.sum(axis=1) * r["omega"] / ft.N
    want = [-float(np.sum(r["occ"][:, None] * np.abs(r["c"][:, :ngw]) ** 2 * r["kpgx"][d][None, :])) for d in range(3)]
    assert np.abs(tot - np.array(want)).max() < 1e-10 * max(1.0, np.abs(want).max())
