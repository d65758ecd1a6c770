"""Subspace dense linear algebra: the PSD/PSDA descent direction (a = c^H Hc, Hc -= c a;
PSDAWavefunctionStepper.cc:65-84, 264-277) and SlaterDet::gram (SlaterDet.cc:1043-1143).
CPU: the plain-C oracle against the golden vectors produced by the reference's own matrix classes (tests/golden/la/) and
against the compiled reference run live.  GPU: the CUDA path through the C ABI against the same vectors and, on seeded
inputs at sizes that cross the GEMM tile boundaries, against the oracle.  Tolerance 1e-10 relative."""
import os

import numpy as np
import pytest

import port as P
import refdrive as R
from util import GOLDEN, TOL, golden_names, load_golden, regen_inputs, relerr, checksum


def _la_golden(name):
    z = np.load(os.path.join(GOLDEN, "la", name + ".npz"))
    return {k: z[k] for k in z.files}


def _inputs(name):
    g = load_golden(name)
    b = P.make_basis(g["cell"], g["ecut"], tuple(g["kpoint"]), bool(g["force_complex"]))
    c, v, occ = regen_inputs(g, b["kpg2"])
    return g, b, c


@pytest.mark.parametrize("name", golden_names("full"))
def test_oracle_la_matches_reference_fixture(name):
    g, b, c = _inputs(name)
    la = _la_golden(name)
    hpsi = g["hpsi"].reshape(c.shape)
    assert abs(checksum(hpsi) - float(la["hpsi_checksum"])) <= 1e-12 * np.abs(hpsi).sum()
    res, a = P.residual(c, hpsi, g["is_real"])
    assert relerr(res, la["resid"]) < TOL
    assert relerr(a, la["resid_a"]) < TOL
    assert relerr(P.gram(c, g["is_real"]), la["gram"]) < TOL


@pytest.mark.skipif(not R.have_ref(), reason="oracle/_ref/ref_driver not built (needs /root/reference)")
@pytest.mark.parametrize("cell,ecut,kpoint,fc,nst", [
    ((10, 0, 0, 0, 10, 0, 0, 0, 10), 6.0, (0, 0, 0), False, 9),
    ((10, 0, 0, 0, 10, 0, 0, 0, 10), 6.0, (0, 0, 0), True, 7),
    ((8, 0, 0, 2.0, 9, 0, -1.0, 0.5, 12), 5.0, (0.5, 0.5, 0.5), False, 5),
])
def test_oracle_la_vs_live_reference(cell, ecut, kpoint, fc, nst):
    r = R.run_reference(R.Case(cell=cell, ecut=ecut, kpoint=kpoint, force_complex=fc, nst=nst), seed=5)
    res, a = P.residual(r["c"], r["hpsi"], r["is_real"])
    assert relerr(res, r["resid"]) < TOL and relerr(a, r["resid_a"]) < TOL
    g = P.gram(r["c"], r["is_real"])
    assert relerr(g, r["gram"]) < TOL
    # the property gram exists for: orthonormal columns (real bases: with the G/-G weighting of the proxy)
    ngw = r["ngw"]
    s = g[:, :ngw].conj() @ g[:, :ngw].T
    if r["is_real"]:
        s = 2.0 * s.real - np.outer(g[:, 0].real, g[:, 0].real)
    assert np.abs(s - np.eye(nst)).max() < 1e-12


def test_oracle_gram_reports_singular_overlap():
    b = P.make_basis((9, 0, 0, 0, 9, 0, 0, 0, 9), 4.0, (0.2, 0, 0), False)
    c = R.synth_coefficients(b["kpg2"], 4.0, 3, b["ngw"], False, 2)
    c[2] = 0.0      # a zero state: the third leading minor is exactly singular
    out = np.array(c, copy=True)
    info = P.lib().qbo_gram(c.shape[1], 3, 0, P._d(out))
    assert info == 3


# ------------------------------------------------------------------------------------------------------------ GPU
def _dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.gpu
@pytest.mark.parametrize("device_ptrs", [True, False])
@pytest.mark.parametrize("name", golden_names("full"))
def test_cuda_la_vs_reference_fixture(name, device_ptrs):
    from qball_b200 import host as H
    g, b, c = _inputs(name)
    la = _la_golden(name)
    wrap = _dev if device_ptrs else (lambda x: np.ascontiguousarray(x).copy())
    back = (lambda t: t.cpu().numpy()) if device_ptrs else (lambda x: x)
    L = H.SubspaceLA(b)
    hc = wrap(g["hpsi"].reshape(c.shape))
    a = wrap(np.zeros((g["nst"], g["nst"]), dtype=np.float64 if g["is_real"] else np.complex128))
    L.residual(wrap(c), hc, a)
    assert relerr(back(hc), la["resid"]) < TOL
    assert relerr(back(a), la["resid_a"]) < TOL
    cg = wrap(c)
    L.gram(cg)
    assert relerr(back(cg), la["gram"]) < TOL
    assert L.launches() > 0


@pytest.mark.gpu
@pytest.mark.parametrize("kpoint,fc,nst,nloc,ldpad,ws", [
    ((0, 0, 0), False, 70, 70, 0, None),          # real basis, one partial tile
    ((0, 0, 0), False, 200, 67, 3, None),         # real basis, band shard of 67 of 200 states, padded rows
    ((0, 0, 0), True, 77, 77, 0, None),           # complex, nst not a multiple of 8
    ((0.25, 0, 0.5), False, 150, 150, 5, None),   # complex k-point, several 64-state tiles, padded rows
    ((0.25, 0, 0.5), False, 150, 40, 0, 1 << 20), # band shard + a workspace that forces plane-wave chunks
])
def test_cuda_la_vs_oracle(kpoint, fc, nst, nloc, ldpad, ws):
    from qball_b200 import host as H
    cell, ecut = (11, 0, 0, 0, 10, 0, 0, 0, 12), 7.0
    b = P.make_basis(cell, ecut, kpoint, fc)
    ldc = b["ngw"] + ldpad
    c = R.synth_coefficients(b["kpg2"], ecut, nst, ldc, b["is_real"], 21)
    # a stand-in for H psi with the same structure (Im = 0 at G = 0 for real bases, zero padding rows)
    hc = R.synth_coefficients(b["kpg2"], ecut, nst, ldc, b["is_real"], 22) * (1.0 + b["kpg2"].max() * 0.1)
    first = (nst - nloc) // 2
    hloc = np.ascontiguousarray(hc[first:first + nloc])
    res_ref, a_ref = P.residual(c, hloc, b["is_real"])
    L = H.SubspaceLA(b)
    if ws:
        L.set_workspace(ws)
    hd = _dev(hloc)
    ad = _dev(np.zeros_like(a_ref))
    L.residual(_dev(c), hd, ad)
    assert relerr(hd.cpu().numpy(), res_ref) < TOL
    assert relerr(ad.cpu().numpy(), a_ref) < TOL
    if ws and not b["is_real"]:
        assert L.query(11) > 1, "expected the sweep to run in several plane-wave chunks"
    assert np.all(hd.cpu().numpy()[:, b["ngw"]:] == 0.0), "padding rows must stay zero"
    # gram on the whole block
    g_ref = P.gram(c, b["is_real"])
    cd = _dev(c)
    L.gram(cd)
    got = cd.cpu().numpy()
    assert relerr(got, g_ref) < TOL
    # a second call on the orthonormal block is the identity to rounding (idempotence)
    L.gram(cd)
    assert relerr(cd.cpu().numpy(), got) < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("kpoint,fc,nst,shards,ws", [((0, 0, 0), False, 11, (4, 4, 3), None), ((0.1, 0.2, 0.3), False, 70, (35, 35), 1 << 20),
                                                      ((0, 0, 0), True, 9, (9,), None), ((0.25, 0, 0), False, 6, (0, 6), None)])
def test_cuda_gram_band_sharded_vs_oracle(kpoint, fc, nst, shards, ws):
    """SlaterDet::gram with the states sharded over ranks (qb200_gram_overlap -> sum over ranks -> qb200_gram_apply), the ranks
    played one after the other on one GPU: every rank's columns must equal those of the oracle's gram of the whole block (and of
    qb200_gram); real and complex bases, uneven shards, an empty shard, a workspace that forces several plane-wave chunks"""
    import torch
    from qball_b200 import host as H
    cell, ecut = (9, 0, 0, 0, 10, 0, 0, 0, 11), 6.0
    b = P.make_basis(cell, ecut, kpoint, fc)
    ngw = b["ngw"]
    c = R.synth_coefficients(b["kpg2"], ecut, nst, ngw + 3, b["is_real"], seed=41)
    want = P.gram(c, b["is_real"])
    la = H.SubspaceLA(b)
    if ws:
        la.set_workspace(ws)
    call = _dev(c)
    S = torch.zeros((nst, nst), dtype=torch.complex128, device="cuda")
    first = 0
    for n in shards:                                  # step 1 on every "rank", step 2 = the sum
        Sr = torch.full((nst, nst), 7.0, dtype=torch.complex128, device="cuda")
        la.gram_overlap(call, first, n, Sr)
        S += Sr
        first += n
    assert first == nst
    if ws:
        assert la.query(11) > 1
    first = 0
    for n in shards:                                  # step 3 on every "rank"
        out = torch.zeros((max(n, 1), ngw + 3), dtype=torch.complex128, device="cuda")
        la.gram_apply(call, S, first, n, out)
        if n:
            got = out.cpu().numpy()[:n]
            assert relerr(got[:, :ngw], want[first:first + n, :ngw]) < TOL
            assert np.all(got[:, ngw:] == 0)
        first += n
    # the one-call form with every state on one rank needs no communicator and may write in place
    c2 = _dev(c)
    la.gram_sharded(None, c2, 0, nst, c2)
    assert relerr(c2.cpu().numpy()[:, :ngw], want[:, :ngw]) < TOL
    la.close()


@pytest.mark.gpu
def test_cuda_gram_singular_overlap_fails_loudly():
    from qball_b200 import capi, host as H
    b = P.make_basis((9, 0, 0, 0, 9, 0, 0, 0, 9), 4.0, (0.2, 0, 0), False)
    c = R.synth_coefficients(b["kpg2"], 4.0, 40, b["ngw"], False, 2)
    c[35] = 0.0
    L = H.SubspaceLA(b)
    cd = _dev(c)
    with pytest.raises(capi.QB200Error):
        L.gram(cd)
    assert np.array_equal(cd.cpu().numpy(), c), "a failed factorisation must leave the block unchanged"


@pytest.mark.gpu
def test_cuda_subspace_la_properties_mgo216_full_size():
    """BASELINE-size block (MgO216: ngw 73447, 768 complex states, 0.9 GB): size-independent properties of the subspace
    algebra -- gram leaves orthonormal columns (and is idempotent on them), a = c^H Hc is returned as computed, and the
    descent direction is orthogonal to the subspace: c^H (Hc - c a) = 0 for orthonormal c."""
    import torch
    from qball_b200 import host as H
    cell, ecut, nst = (23.1, 0, 0, 0, 23.1, 0, 0, 0, 23.1), 25.0, 768
    b = P.make_basis(cell, ecut, (0, 0, 0), True)
    assert b["ngw"] == 73447
    c = torch.from_numpy(R.synth_coefficients(b["kpg2"], ecut, nst, b["ngw"], False, seed=41)).cuda()
    hc = torch.from_numpy(R.synth_coefficients(b["kpg2"], ecut, nst, b["ngw"], False, seed=42)).cuda()
    L = H.SubspaceLA(b)
    L.gram(c)
    eye = torch.eye(nst, dtype=torch.complex128, device="cuda")
    s = c.conj() @ c.T                                       # (nst, ldc) blocks: s[m, n] = <c_m | c_n>
    assert float((s - eye).abs().max()) < 1e-12
    c2 = c.clone()
    L.gram(c2)
    assert float((c2 - c).abs().max()) < 1e-12 * float(c.abs().max())
    a_want = (c.conj() @ hc.T).T.contiguous()                # row n = column n of a: a[m, n] = <c_m | hc_n>
    a = torch.zeros((nst, nst), dtype=torch.complex128, device="cuda")
    res = hc.clone()
    L.residual(c, res, a)
    scale = float(a_want.abs().max())
    assert float((a - a_want).abs().max()) < 1e-11 * scale
    assert float((c.conj() @ res.T).abs().max()) < 1e-11 * scale
    # band-sharded form: the columns of a middle shard equal the same columns of the full result
    lo, hi = 200, 331
    part = hc[lo:hi].clone()
    L.residual(c, part)
    assert float((part - res[lo:hi]).abs().max()) < 1e-12 * float(res.abs().max())


@pytest.mark.gpu
@pytest.mark.parametrize("kpoint,fc", [((0, 0, 0), False), ((0.1, 0.2, 0.3), False), ((0, 0, 0), True)])
def test_cuda_psda_update_vs_oracle(kpoint, fc):
    """the rest of PSDAWavefunctionStepper::update (PSDAWavefunctionStepper.cc:93-225, 281-395) with Preconditioner::apply
    (Preconditioner.cc:118-139): first call without extrapolation, second with Anderson's theta, a third with a direction
    that drives theta into the clipped range; every block (c, dc, c_last, dc_last) against the oracle's restatement"""
    import torch
    from qball_b200 import host as H
    cell, ecut, nst = (10, 0, 0, 0, 11, 0, 0, 0, 12), 6.0, 9
    b = P.make_basis(cell, ecut, kpoint, fc)
    ngw, ldc = b["ngw"], b["ngw"] + 3
    occ = R.synth_occ(nst, nst - 2)
    prec = P.preconditioner_diag(b["kpg2"], 2.0)
    la = H.SubspaceLA(b)
    blocks = [R.synth_coefficients(b["kpg2"], ecut, nst, ldc, b["is_real"], seed=s) for s in (91, 92)]
    c_ref, cl_ref, dl_ref = blocks[0].copy(), np.zeros_like(blocks[0]), np.zeros_like(blocks[0])
    cd, cld, dld = (torch.from_numpy(a.copy()).cuda() for a in (c_ref, cl_ref, dl_ref))
    for it, (seed, scale) in enumerate([(93, 0.3), (94, 0.25), (95, -4.0)]):
        dc = scale * R.synth_coefficients(b["kpg2"], ecut, nst, ldc, b["is_real"], seed=seed)
        dcd = torch.from_numpy(dc.copy()).cuda()
        th_ref = P.psda_update(ngw, b["is_real"], c_ref, dc, cl_ref, dl_ref, occ, prec, it > 0)
        th = la.psda_update(cd, dcd, cld, dld, occ, prec if it % 2 == 0 else torch.from_numpy(prec).cuda(), it > 0)
        assert abs(th - th_ref) <= 1e-11 * max(1.0, abs(th_ref)), (it, th, th_ref)
        for got, want in ((cd, c_ref), (dcd, dc), (cld, cl_ref), (dld, dl_ref)):
            assert relerr(got.cpu().numpy(), want) < 1e-12
        assert np.all(cd.cpu().numpy()[:, ngw:] == c_ref[:, ngw:])


@pytest.mark.gpu
def test_cuda_scf_iterations_stay_on_device_vs_oracle():
    """three electronic iterations as BOSampleStepper drives them with wf_dyn PSDA -- H psi -> a = c^H Hc, Hc -= c a ->
    preconditioner + Anderson update -> SlaterDet::gram -> density -- with the block resident on the device throughout
    (only v goes in, rho / E_nl / E_kin / theta come out), against the same sequence over the oracle"""
    import torch
    from qball_b200 import host as H
    cell, ecut, nst = (10, 0, 0, 0, 11, 0, 0, 0, 12), 6.0, 6
    b = P.make_basis(cell, ecut, (0, 0, 0), False)
    grid = P.density_grid(cell, ecut)
    ngw = b["ngw"]
    N = grid[0] * grid[1] * grid[2]
    occ = R.synth_occ(nst, nst)
    v = 0.2 * R.synth_potential(*grid, seed=97)
    prec = P.preconditioner_diag(b["kpg2"], 2.0)
    rng = np.random.default_rng(98)
    species = [dict(na=2, npr=4, lproj=np.array([0, 1, 1, 1], dtype=np.int32), wt=np.array([0.9, -0.4, -0.4, -0.4]),
                    twnl=rng.standard_normal((4, ngw)) * np.exp(-b["kpg2"] / 4.0)[None, :], tau=rng.uniform(0, 10, (2, 3)))]
    c0 = R.synth_coefficients(b["kpg2"], ecut, nst, ngw, True, seed=96)
    c0 = P.gram(c0, True)                                                 # orthonormal start (SlaterDet::gram)
    # oracle sequence
    oft = P.FT(b, *grid)
    c, cl, dl = c0.copy(), np.zeros_like(c0), np.zeros_like(c0)
    ref = []
    for it in range(3):
        enl, hc = P.hpsi(b, oft, np.ascontiguousarray(c), v, occ, species)
        ekin = P.ekin_sums(b["kpg2"], c, occ, True)[0][0]
        hc, _ = P.residual(np.ascontiguousarray(c), hc, True)
        th = P.psda_update(ngw, True, c, hc, cl, dl, occ, prec, it > 0)
        c = P.gram(c, True)
        rho = oft.compute_density(np.ascontiguousarray(c), occ / b["omega"], np.zeros(N))
        ref.append((enl, ekin, th, rho.copy(), c.copy()))
    # device sequence: the coefficient block is uploaded once and never read back inside the loop
    ft, nlp, la = H.FourierTransform(b, *grid), H.NonLocalPotential(b, species), H.SubspaceLA(b)
    cd = torch.from_numpy(c0.copy()).cuda()
    cld, dld, hd = torch.zeros_like(cd), torch.zeros_like(cd), torch.zeros_like(cd)
    kd = torch.from_numpy(b["kpg2"]).cuda()
    for it in range(3):
        enl = H.hpsi(ft, nlp, cd, occ, v, kd, hd)                           # v: host array in
        ekin = H.ekin_sums(ft, cd, occ, True, kd)[0]
        la.residual(cd, hd)
        th = la.psda_update(cd, hd, cld, dld, occ, prec, it > 0)
        la.gram(cd)
        rho = np.zeros(N)
        H.compute_density(ft, cd, 1.0, occ, b["omega"], rho)                # rho: host array out
        e_ref, k_ref, t_ref, r_ref, c_ref = ref[it]
        assert abs(enl - e_ref) < 1e-10 * max(1.0, abs(e_ref)) and abs(ekin - k_ref) < 1e-10 * abs(k_ref)
        assert abs(th - t_ref) < 1e-8 * max(1.0, abs(t_ref)), (it, th, t_ref)
        assert relerr(rho, r_ref) < TOL and relerr(cd.cpu().numpy(), c_ref) < TOL


@pytest.mark.parametrize("name", golden_names("full"))
def test_oracle_diag_matches_reference_fixture(name):
    """eigenvalues of psi^H (H psi): the oracle (LAPACK via numpy on the reference's formula) against the reference's own
    syevd / heev('l') calls (tests/golden/la/*.npz diag_w, Wavefunction.cc:1538-1539, 1612, 1641, 1693)"""
    g, b, c = _inputs(name)
    la = _la_golden(name)
    hpsi = g["hpsi"].reshape(c.shape)
    w, _, _ = P.diag(c, hpsi, g["is_real"])
    assert np.abs(w - la["diag_w"]).max() <= 1e-12 * max(1.0, np.abs(la["diag_w"]).max())


@pytest.mark.gpu
@pytest.mark.parametrize("name", golden_names("full"))
def test_cuda_diag_vs_reference_fixture(name):
    import torch
    from qball_b200 import host as H
    g, b, c = _inputs(name)
    la_g = _la_golden(name)
    hpsi = np.ascontiguousarray(g["hpsi"].reshape(c.shape))
    la = H.SubspaceLA(b)
    w, sweeps = la.diag(c.copy(), hpsi, eigvec=False)                       # host pointers
    assert np.abs(w - la_g["diag_w"]).max() <= 1e-10 * max(1.0, np.abs(la_g["diag_w"]).max()), (w, la_g["diag_w"])
    assert 1 <= sweeps < 30


@pytest.mark.gpu
@pytest.mark.parametrize("blocked", ["1", "0"])
def test_cuda_diag_sweeps_from_a_cuda_graph_on_a_user_stream(blocked, monkeypatch):
    """on a capturable stream the launches of a Jacobi sweep (blocked: 3 per block step; element-wise: 2 per rotation step) are
    captured once and replayed (the legacy default stream launches directly): same eigenvalues and rotated states"""
    import torch
    from qball_b200 import host as H
    monkeypatch.setenv("QB200_JACOBI_BLOCK", blocked)
    cell, ecut, nst = (10, 0, 0, 0, 9, 0, 0, 0, 11), 6.0, 70
    b = P.make_basis(cell, ecut, (0.2, 0.1, 0.0), False)
    c = P.gram(R.synth_coefficients(b["kpg2"], ecut, nst, b["ngw"], False, 61), False)
    hc = R.synth_coefficients(b["kpg2"], ecut, nst, b["ngw"], False, 62) + c * (1.0 + b["kpg2"])[None, :]
    res = {}
    for mode in ("graph", "direct"):
        st = torch.cuda.Stream() if mode == "graph" else None
        la = H.SubspaceLA(b, stream=st)
        with torch.cuda.stream(st) if st is not None else torch.cuda.stream(torch.cuda.current_stream()):
            cd, hd = _dev(c), _dev(hc)
            torch.cuda.synchronize()
            w, _ = la.diag(cd, hd)
            torch.cuda.synchronize()
        assert la.query(13) == (1 if mode == "graph" else 0)
        res[mode] = (np.asarray(w).copy(), cd.cpu().numpy().copy())
        la.close()
    assert np.allclose(res["graph"][0], res["direct"][0], rtol=0, atol=1e-12 * np.abs(res["direct"][0]).max())
    assert relerr(res["graph"][1], res["direct"][1]) < 1e-10


@pytest.mark.gpu
@pytest.mark.parametrize("kpoint,nst", [((0, 0, 0), 5), ((0.2, 0.1, 0.0), 64), ((0.2, 0.1, 0.0), 97), ((0, 0, 0), 200)])
def test_cuda_diag_blocked_and_cyclic_jacobi_agree(kpoint, nst, monkeypatch):
    """the blocked Jacobi method (default: 64 x 64 pivots diagonalised in shared memory, block rotations applied as small GEMMs)
    against the element-wise cyclic Jacobi and against numpy's eigvalsh of the same subspace matrix: orders below one block,
    exactly one pivot, an odd number of blocks (padding block), real and complex bases"""
    from qball_b200 import host as H
    cell, ecut = (10, 0, 0, 0, 9, 0, 0, 0, 11), 7.0
    b = P.make_basis(cell, ecut, kpoint, False)
    c = P.gram(R.synth_coefficients(b["kpg2"], ecut, nst, b["ngw"], b["is_real"], 71), b["is_real"])
    hc = R.synth_coefficients(b["kpg2"], ecut, nst, b["ngw"], b["is_real"], 72) + c * (1.0 + b["kpg2"])[None, :]
    w_ref = np.linalg.eigvalsh(np.tril(P.subspace_h(c, hc, b["is_real"])) + np.tril(P.subspace_h(c, hc, b["is_real"]), -1).conj().T)
    res = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("QB200_JACOBI_BLOCK", mode)
        la = H.SubspaceLA(b)
        cd, hd = _dev(c), _dev(hc)
        w, sweeps = la.diag(cd, hd)
        assert la.query(14) == int(mode) and 1 <= sweeps < 30
        scale = np.abs(w_ref).max()
        assert np.abs(np.asarray(w) - w_ref).max() < 1e-12 * scale
        # the rotated states diagonalise H in the subspace: c'^H (H c') = diag(w), with H c' = (H c) z
        res[mode] = (np.asarray(w).copy(), cd.cpu().numpy().copy())
        la.close()
    # eigenvectors agree up to phases where the eigenvalues are well separated: compare the projectors c' c'^H through |<a|b>|
    gap = np.min(np.diff(res["1"][0])) / np.abs(w_ref).max()
    if gap > 1e-6:
        a_, b_ = res["1"][1][:, :b["ngw"]], res["0"][1][:, :b["ngw"]]
        if b["is_real"]:
            ov = np.abs(2.0 * np.einsum("ng,ng->n", np.conj(a_), b_).real - a_[:, 0].real * b_[:, 0].real)
        else:
            ov = np.abs(np.einsum("ng,ng->n", np.conj(a_), b_))
        assert np.abs(ov - 1.0).max() < 1e-8


@pytest.mark.gpu
@pytest.mark.parametrize("kpoint,fc,nst", [((0, 0, 0), False, 37), ((0.1, 0.2, 0.3), False, 130), ((0, 0, 0), True, 64)])
def test_cuda_diag_eigenpairs_vs_oracle(kpoint, fc, nst):
    """Wavefunction::diag with eigenvectors at sizes that cross the GEMM tiles (odd n included): eigenvalues against LAPACK
    (the oracle) to 1e-10; the rotated block c z: z unitary, z^H h z = diag(w), and the subspace unchanged"""
    import torch
    from qball_b200 import host as H
    cell, ecut = (10, 0, 0, 0, 11, 0, 0, 0, 12), 6.0
    b = P.make_basis(cell, ecut, kpoint, fc)
    ngw, ldc = b["ngw"], b["ngw"] + 2
    c = P.gram(R.synth_coefficients(b["kpg2"], ecut, nst, ldc, b["is_real"], seed=71), b["is_real"])
    hc = R.synth_coefficients(b["kpg2"], ecut, nst, ldc, b["is_real"], seed=72)
    # a Hermitian "H c": symmetrise the subspace matrix through hc <- hc + c (h^H - h)/2 is not needed: diag reads the LOWER
    # triangle only (LAPACK 'l'), so any hc defines a Hermitian problem -- exactly what the reference solves
    w_ref, z_ref, h = P.diag(c, hc, b["is_real"])
    la = H.SubspaceLA(b)
    cd = torch.from_numpy(c.copy()).cuda()
    w, sweeps = la.diag(cd, torch.from_numpy(hc).cuda(), eigvec=True)
    assert np.abs(w - w_ref).max() <= 1e-10 * np.abs(w_ref).max(), np.abs(w - w_ref).max()
    cn = cd.cpu().numpy()
    assert np.all(cn[:, ngw:] == 0)
    # z implied by c_new = c z (c orthonormal): z = S(c, c_new), with the reference's real-basis inner product
    z = P.subspace_h(c, cn, b["is_real"])
    hl = np.tril(h, -1) + np.tril(h, -1).conj().T + np.diag(np.diag(h).real)   # the matrix LAPACK 'l' sees (real diagonal)
    n = nst
    assert np.abs(z.conj().T @ z - np.eye(n)).max() < 1e-11
    assert np.abs(z.conj().T @ hl @ z - np.diag(w)).max() < 1e-10 * np.abs(w_ref).max()
    # eigenvalues only: the block is left untouched
    cd2 = torch.from_numpy(c.copy()).cuda()
    w2, _ = la.diag(cd2, torch.from_numpy(hc).cuda(), eigvec=False)
    assert np.array_equal(cd2.cpu().numpy(), c) and np.abs(w2 - w).max() <= 1e-13 * np.abs(w).max()
