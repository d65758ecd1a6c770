"""GPU: the CUDA path, called through the C ABI (qball_b200.host -> libqball_b200.so), against
  (1) the committed golden vectors produced by the reference itself (tests/golden/*.npz), and
  (2) the plain-C oracle on seeded inputs for shapes/paths the fixtures do not cover.
Tolerance: 1e-10 relative (BASELINE.json north_star) -- FP64 FFT/GEMM results are not bit-identical between
algorithms; the index tables (integer work) are checked exactly in the CPU tests."""
import numpy as np
import pytest
import torch

import port as P
import refdrive as R
from qball_b200 import host as H
from util import TOL, compare, golden_names, load_golden, regen_inputs, relerr

pytestmark = pytest.mark.gpu


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _run_fixture(name, use_device_ptrs, force_split, monkeypatch, use_lattice=True, anl_bytes=None):
    if force_split:
        monkeypatch.setenv("QB200_FORCE_SPLIT", "1")
    g = load_golden(name)
    b = P.make_basis(g["cell"], g["ecut"], tuple(g["kpoint"]), bool(g["force_complex"]))
    c, v, occ = regen_inputs(g, b["kpg2"])
    ngw, N = g["ngw"], g["np0"] * g["np1"] * g["np2"]
    ft = H.FourierTransform(b, g["np0"], g["np1"], g["np2"])
    assert ft.fused() == (not force_split)
    assert ft.nvec() == (2 * g["nrods"] - 1 if g["is_real"] else g["nrods"])
    wrap = _dev if use_device_ptrs else (lambda a: np.ascontiguousarray(a))
    back = (lambda t: t.cpu().numpy()) if use_device_ptrs else (lambda a: a)
    # single transforms
    f = wrap(np.zeros(N, dtype=np.complex128))
    ft.backward(wrap(c[0, :ngw].copy()), f)
    compare(g, "bwd0", back(f))
    fv = wrap(back(f) * v)
    cc = wrap(np.zeros(ngw, dtype=np.complex128))
    ft.forward(fv, cc)
    compare(g, "fwd0", back(cc))
    if g["is_real"] and g["nst"] >= 2:
        ft.backward(wrap(c[0, :ngw].copy()), f, c2=wrap(c[1, :ngw].copy()))
        compare(g, "bwdpair01", back(f))
        fv = wrap(back(f) * v)
        c1, c2 = wrap(np.zeros(ngw, dtype=np.complex128)), wrap(np.zeros(ngw, dtype=np.complex128))
        ft.forward(fv, c1, c2)
        compare(g, "fwdpair0", back(c1))
        compare(g, "fwdpair1", back(c2))
    # rs_mul_add, density, nonlocal, whole H psi
    cw, vw = wrap(c), wrap(v)
    cp = wrap(np.zeros_like(c))
    H.rs_mul_add(ft, cw, vw, cp)
    compare(g, "hloc", back(cp))
    rho = wrap(np.zeros(N))
    H.compute_density(ft, cw, 1.0, occ, g["omega"], rho)
    compare(g, "rho", back(rho))
    nlp = H.NonLocalPotential(b, g["species"], use_lattice=use_lattice)
    if anl_bytes is not None:
        nlp.set_workspace(anl_bytes)
    cp = wrap(np.zeros_like(c))
    enl = nlp.energy(cw, occ, True, cp)
    assert abs(enl - g["enl"]) <= 1e-10 * max(1.0, abs(g["enl"])), (enl, g["enl"])
    if g["nsp"]:
        compare(g, "hnl", back(cp))
    out = wrap(np.zeros_like(c))
    enl2 = H.hpsi(ft, nlp, cw, occ, vw, wrap(b["kpg2"]), out)
    assert abs(enl2 - g["enl"]) <= 1e-10 * max(1.0, abs(g["enl"]))
    compare(g, "hpsi", back(out))
    if anl_bytes is not None and g["nsp"] and g["ngw"] > 512:
        assert nlp.query(11) > 1, "expected the projector sweep to run in several plane-wave chunks"


@pytest.mark.parametrize("name", golden_names())
def test_cuda_vs_reference_fixture_device_pointers(name, monkeypatch):
    _run_fixture(name, True, False, monkeypatch)


@pytest.mark.parametrize("name", golden_names("full"))
def test_cuda_vs_reference_fixture_host_pointers(name, monkeypatch):
    _run_fixture(name, False, False, monkeypatch)


@pytest.mark.parametrize("name", golden_names("full"))
def test_cuda_projectors_chunked_sweep(name, monkeypatch):
    """a workspace too small for the whole anl block: the two GEMM sweeps run chunk by chunk (the Au992 regime)"""
    _run_fixture(name, True, False, monkeypatch, anl_bytes=1 << 20)


@pytest.mark.parametrize("name", golden_names())
def test_cuda_projectors_sincos_path(name, monkeypatch):
    """without qb200_nl_set_lattice the projector tiles evaluate exp(-i (k+G).tau) with sincos per (atom, G)"""
    _run_fixture(name, True, False, monkeypatch, use_lattice=False)


def test_cuda_mgo216_compiled_shape_and_generic_kernel(monkeypatch):
    """the MgO216 grid runs the shape-specialised plane kernel; the generic kernel must give the same answers there"""
    g = load_golden("mgo216_shape_112cubed")
    b = P.make_basis(g["cell"], g["ecut"], tuple(g["kpoint"]), bool(g["force_complex"]))
    ft = H.FourierTransform(b, g["np0"], g["np1"], g["np2"])
    assert ft.query(10) == 2, "MgO216 plan did not select the compiled shape (one warp per column block)"
    assert ft.query(17) == 1, "MgO216 plan did not select the tensor-memory plane kernel (k_plane_t)"
    del ft
    _run_fixture("mgo216_shape_112cubed", True, False, monkeypatch)
    # the shared-memory-only kernel of the same compiled shape (k_plane_s)
    monkeypatch.setenv("QB200_PLANE_T", "0")
    ft = H.FourierTransform(b, g["np0"], g["np1"], g["np2"])
    assert ft.query(10) == 2 and ft.query(17) == 0
    del ft
    _run_fixture("mgo216_shape_112cubed", True, False, monkeypatch)
    monkeypatch.delenv("QB200_PLANE_T")
    # opt-in variants of k_plane_t: the density with its rho plane in shared memory (k_plane_td), the other warp geometries
    for var, val in (("QB200_T_DENS_SMEM", "1"), ("QB200_T_DENS", "1"), ("QB200_T_HPSI", "1")):
        monkeypatch.setenv(var, val)
        _run_fixture("mgo216_shape_112cubed", True, False, monkeypatch)
        monkeypatch.delenv(var)
    # the z-column kernels: tensor-memory form by default (k_zcol_bwd_t / k_zcol_fwd_t), the shared-memory tiles (v2) otherwise
    ft = H.FourierTransform(b, g["np0"], g["np1"], g["np2"])
    assert ft.query(18) == 1, "MgO216 plan did not select the tensor-memory z-column kernels"
    del ft
    monkeypatch.setenv("QB200_ZCOL_T", "0")
    ft = H.FourierTransform(b, g["np0"], g["np1"], g["np2"])
    assert ft.query(18) == 0 and ft.query(16) == 1
    del ft
    _run_fixture("mgo216_shape_112cubed", True, False, monkeypatch)
    monkeypatch.delenv("QB200_ZCOL_T")
    monkeypatch.setenv("QB200_NO_STATIC", "1")
    ft = H.FourierTransform(b, g["np0"], g["np1"], g["np2"])
    assert ft.query(10) == 0
    del ft
    _run_fixture("mgo216_shape_112cubed", True, False, monkeypatch)
    monkeypatch.setenv("QB200_NO_STAGE", "1")
    _run_fixture("mgo216_shape_112cubed", True, False, monkeypatch)
    # the other compiled geometry of the same shape (7 groups of 64 threads instead of one warp per column block)
    monkeypatch.delenv("QB200_NO_STATIC")
    monkeypatch.delenv("QB200_NO_STAGE")
    monkeypatch.setenv("QB200_GROUP_THREADS", "64")
    ft = H.FourierTransform(b, g["np0"], g["np1"], g["np2"])
    assert ft.query(10) == 1
    del ft
    _run_fixture("mgo216_shape_112cubed", True, False, monkeypatch)
    # ... and the warp-owned x phase (k_plane_w, opt-in)
    monkeypatch.delenv("QB200_GROUP_THREADS")
    monkeypatch.setenv("QB200_PLANE_W", "1")
    ft = H.FourierTransform(b, g["np0"], g["np1"], g["np2"])
    assert ft.query(10) == 3
    del ft
    _run_fixture("mgo216_shape_112cubed", True, False, monkeypatch)


@pytest.mark.parametrize("name", ["gamma_triclinic_si_h", "kpoint_cubic_au_oncv", "sih4_60cubed"])
def test_cuda_split_path_vs_reference_fixture(name, monkeypatch):
    """the split (large-plane) kernels, forced on small grids"""
    _run_fixture(name, True, True, monkeypatch)


@pytest.mark.parametrize("cell,ecut,kpoint,fc,nst,ldpad", [
    ((11, 0, 0, 0, 12, 0, 0, 0, 13), 6.0, (0, 0, 0), False, 7, 5),         # Gamma, odd tail, padded ldc, 3 pair batches
    ((11, 0, 0, 0, 12, 0, 0, 0, 13), 6.0, (0.1, 0.2, 0.3), False, 5, 3),  # complex, padded ldc
    ((16, 0, 0, 0, 9, 0, 0, 0, 9), 9.0, (0, 0, 0.5), False, 3, 0),         # radix mix incl. 5/3/9
])
def test_cuda_vs_oracle_seeded(cell, ecut, kpoint, fc, nst, ldpad):
    b = P.make_basis(cell, ecut, kpoint, fc)
    grid = P.density_grid(cell, ecut)
    ldc = b["ngw"] + ldpad
    c = R.synth_coefficients(b["kpg2"], ecut, nst, ldc, b["is_real"], seed=5)
    v = R.synth_potential(*grid, seed=9)
    occ = R.synth_occ(nst, nst - 1)
    oft = P.FT(b, *grid)
    ft = H.FourierTransform(b, *grid)
    ft.set_workspace(3 * ft.nvec() * grid[2] * 16)  # 3 units per batch -> several batches
    want = oft.rs_mul_add(c, v, np.zeros_like(c))
    P.kinetic_add(b["kpg2"], c, want)
    got = _dev(np.zeros_like(c))
    H.rs_mul_add(ft, _dev(c), _dev(v), got, kpg2=_dev(b["kpg2"]))
    got = got.cpu().numpy()
    assert relerr(got[:, :b["ngw"]], want[:, :b["ngw"]]) < TOL
    assert np.all(got[:, b["ngw"]:] == 0)  # padding rows untouched (SlaterDet.cc:2784-2787)
    rho_want = oft.compute_density(c, occ / b["omega"], np.full(oft.N, 0.25))
    rho = _dev(np.full(oft.N, 0.25))
    H.compute_density(ft, _dev(c), 1.0, occ, b["omega"], rho)
    assert relerr(rho.cpu().numpy(), rho_want) < TOL


@pytest.mark.parametrize("fc,nst,ldpad", [(True, 7, 3), (False, 5, 0)])
def test_cuda_mgo216_shape_vs_oracle_seeded(fc, nst, ldpad):
    """the benchmark's grid (MgO216 cell, 50 Ry, 112^3) through the tensor-memory kernels against the oracle on seeded inputs:
    complex states with a padded leading dimension, zero occupations (units the density kernel skips) and a workspace of
    three units per batch; and the Gamma-point REAL basis on the same grid (pairs through k_plane_t, the odd tail with its
    imaginary part dropped through k_plane_s, z columns through the shared-memory tiles)"""
    cell, ecut = (23.1, 0, 0, 0, 23.1, 0, 0, 0, 23.1), 25.0
    b = P.make_basis(cell, ecut, (0, 0, 0), fc)
    grid = P.density_grid(cell, ecut)
    assert grid == (112, 112, 112) and b["is_real"] == (not fc)
    ldc = b["ngw"] + ldpad
    c = R.synth_coefficients(b["kpg2"], ecut, nst, ldc, b["is_real"], seed=15)
    v = R.synth_potential(*grid, seed=19)
    occ = R.synth_occ(nst, nst - 1)
    occ[1] = 0.0
    oft = P.FT(b, *grid)
    ft = H.FourierTransform(b, *grid)
    assert ft.query(17) == 1 and ft.query(18) == (1 if fc else 0)
    ft.set_workspace(3 * ft.nvec() * grid[2] * 16)
    want = oft.rs_mul_add(c, v, np.zeros_like(c))
    P.kinetic_add(b["kpg2"], c, want)
    got = _dev(np.zeros_like(c))
    H.rs_mul_add(ft, _dev(c), _dev(v), got, kpg2=_dev(b["kpg2"]))
    got = got.cpu().numpy()
    assert relerr(got[:, :b["ngw"]], want[:, :b["ngw"]]) < TOL
    assert np.all(got[:, b["ngw"]:] == 0)
    rho_want = oft.compute_density(c, occ / b["omega"], np.full(oft.N, 0.25))
    rho = _dev(np.full(oft.N, 0.25))
    H.compute_density(ft, _dev(c), 1.0, occ, b["omega"], rho)
    assert relerr(rho.cpu().numpy(), rho_want) < TOL
    # the same calls with host pointers (sliced, pipelined uploads)
    goth = np.zeros_like(c)
    H.rs_mul_add(ft, c, v, goth, kpg2=b["kpg2"])
    assert relerr(goth[:, :b["ngw"]], want[:, :b["ngw"]]) < TOL
    rhoh = np.full(oft.N, 0.25)
    H.compute_density(ft, c, 1.0, occ, b["omega"], rhoh)
    assert relerr(rhoh, rho_want) < TOL


def test_properties_mgo216_full_size():
    """BASELINE-size shape (MgO216: 112^3, ngw 73447): size-independent properties -- fwd(bwd(c)) = c, Parseval,
    linearity of H_loc, integral of rho = sum of occupations times norms."""
    cell, ecut = (23.1, 0, 0, 0, 23.1, 0, 0, 0, 23.1), 25.0
    b = P.make_basis(cell, ecut, (0, 0, 0), True)
    assert b["ngw"] == 73447 and b["nrods"] == 2109
    grid = P.density_grid(cell, ecut)
    assert grid == (112, 112, 112)
    N = 112 ** 3
    nst = 6
    c = R.synth_coefficients(b["kpg2"], ecut, nst, b["ngw"], False, seed=21)
    v = R.synth_potential(*grid, seed=22)
    ft = H.FourierTransform(b, *grid)
    cd = _dev(c)
    f = torch.zeros(N, dtype=torch.complex128, device="cuda")
    ft.backward(cd[0].contiguous(), f)
    nrm = float((f.abs() ** 2).sum().item()) / N
    assert abs(nrm - float(np.vdot(c[0], c[0]).real)) < 1e-10 * nrm       # Parseval
    back = torch.zeros(b["ngw"], dtype=torch.complex128, device="cuda")
    ft.forward(f, back)
    assert relerr(back.cpu().numpy(), c[0]) < 1e-12                        # round trip
    vd = _dev(v)
    h = torch.zeros_like(cd)
    H.rs_mul_add(ft, cd, vd, h)
    lin = torch.zeros(1, b["ngw"], dtype=torch.complex128, device="cuda")
    comb = (0.3 * cd[1] - 1.7j * cd[2]).reshape(1, -1).contiguous()
    H.rs_mul_add(ft, comb, vd, lin)
    assert relerr(lin[0].cpu().numpy(), (0.3 * h[1] - 1.7j * h[2]).cpu().numpy()) < TOL   # linearity
    # <c_m| V c_n> is Hermitian for real v
    m = (cd.conj() @ h.T).cpu().numpy()
    assert np.abs(m - m.conj().T).max() < 1e-10 * np.abs(m).max()
    occ = np.array([2, 2, 1.5, 0, 1, 0.25])
    rho = torch.zeros(N, dtype=torch.float64, device="cuda")
    H.compute_density(ft, cd, 1.0, occ, b["omega"], rho)
    nel = float(rho.sum().item()) * b["omega"] / N                          # ChargeDensity.cc:525
    want = float(sum(occ[n] * np.vdot(c[n], c[n]).real for n in range(nst)))
    assert abs(nel - want) < 1e-10 * want
    assert float(rho.min().item()) >= 0.0


def test_no_device_fallback_is_an_error():
    from qball_b200 import capi
    L = capi.load()
    import ctypes as C
    h = C.c_void_p()
    a = (C.c_int * 1)(0)
    s = (C.c_int * 1)(3)
    rc = L.qb200_plan_create(C.byref(h), 99, 8, 8, 8, 1, a, a, a, s, 0, 0, 0)
    assert rc == -2 and b"device" in L.qb200_last_error()


@pytest.mark.parametrize("kpoint,nst", [((0, 0, 0), 259), ((0.1, 0.2, 0.3), 300)])
def test_cuda_host_blocks_pipelined_and_resident(kpoint, nst):
    """HOST coefficient blocks: qb200_hpsi / qb200_compute_density pipeline them in blocks of 128 states (upload,
    compute, download overlapped); results must equal the device-pointer path and the oracle, for a real basis with an
    odd tail in the last block and a complex one.  With a coefficient tag the second call reuses the device copy."""
    cell, ecut = (11, 0, 0, 0, 12, 0, 0, 0, 13), 6.0
    b = P.make_basis(cell, ecut, kpoint, False)
    grid = P.density_grid(cell, ecut)
    ngw = b["ngw"]
    c = R.synth_coefficients(b["kpg2"], ecut, nst, ngw + 2, b["is_real"], seed=31)
    v = R.synth_potential(*grid, seed=32)
    occ = R.synth_occ(nst, nst - 7)
    rng = np.random.default_rng(33)
    species = [dict(na=5, npr=4, lproj=np.array([0, 1, 1, 1], dtype=np.int32), wt=np.array([1.3, -0.6, -0.6, -0.6]),
                    twnl=rng.standard_normal((4, ngw)) * np.exp(-b["kpg2"] / 4.0)[None, :], tau=rng.uniform(0, 11, (5, 3)))]
    ft = H.FourierTransform(b, *grid)
    nlp = H.NonLocalPotential(b, species)
    # device pointers: one block
    out_d = _dev(np.zeros_like(c))
    enl_d = H.hpsi(ft, nlp, _dev(c), occ, _dev(v), _dev(b["kpg2"]), out_d)
    rho_d = _dev(np.zeros(grid[0] * grid[1] * grid[2]))
    H.compute_density(ft, _dev(c), 1.0, occ, b["omega"], rho_d)
    # host pointers (pinned): blocks of 128 states
    hc = torch.from_numpy(c.copy()).pin_memory()
    hout = torch.zeros_like(hc).pin_memory()
    hrho = torch.zeros(grid[0] * grid[1] * grid[2], dtype=torch.float64).pin_memory()
    ft.set_coefficient_tag(1)
    enl_h = H.hpsi(ft, nlp, hc, occ, v, b["kpg2"], hout)
    H.compute_density(ft, hc, 1.0, occ, b["omega"], hrho)          # same tag: no second upload
    assert abs(enl_h - enl_d) <= 1e-12 * max(1.0, abs(enl_d))
    assert relerr(hout.numpy(), out_d.cpu().numpy()) < 1e-12
    assert relerr(hrho.numpy(), rho_d.cpu().numpy()) < 1e-12
    assert np.all(hout.numpy()[:, ngw:] == 0)
    # oracle on a few states of the first and the last block
    sel = [0, 1, nst - 2, nst - 1] if not b["is_real"] else [0, 1]
    oft = P.FT(b, *grid)
    _, h_ref = P.hpsi(b, oft, np.ascontiguousarray(c[sel]), v, occ[sel], species)
    assert relerr(hout.numpy()[sel][:, :ngw], h_ref[:, :ngw]) < TOL
    # residency semantics: same tag -> the device copy is used; a new tag -> the host block is uploaded again
    hc[:, :ngw] *= 2.0
    hrho.zero_()
    H.compute_density(ft, hc, 1.0, occ, b["omega"], hrho)
    assert relerr(hrho.numpy(), rho_d.cpu().numpy()) < 1e-12
    ft.set_coefficient_tag(2)
    hrho.zero_()
    H.compute_density(ft, hc, 1.0, occ, b["omega"], hrho)
    assert relerr(hrho.numpy(), 4.0 * rho_d.cpu().numpy()) < 1e-12
    ft.set_coefficient_tag(0)
    hc[:, :ngw] *= 0.5
    hrho.zero_()
    H.compute_density(ft, hc, 1.0, occ, b["omega"], hrho)        # tag 0: always uploaded
    assert relerr(hrho.numpy(), rho_d.cpu().numpy()) < 1e-12


def test_properties_au992_full_size_split_path(monkeypatch):
    """BASELINE-size shape (gold benchmark: 252 x 252 x 896 grid, ngw 2.84 M, planes larger than shared memory): the
    compiled-shape split kernels against the run-time-shape split kernels (which the forced-split fixtures tie to the
    reference), plus size-independent properties: Parseval, fwd(bwd(c)) = c, integral of rho."""
    cell, ecut, kpoint = (30.84, 0, 0, 0, 30.84, 0, 0, 0, 119.505), 65.0, (1e-7, 0, 0)
    b = P.make_basis(cell, ecut, kpoint, False)
    grid = P.density_grid(cell, ecut)
    assert grid == (252, 252, 896) and not b["is_real"]
    N, ngw, nst = grid[0] * grid[1] * grid[2], b["ngw"], 3
    c = R.synth_coefficients(b["kpg2"], ecut, nst, ngw, False, seed=41)
    v = R.synth_potential(*grid, seed=42)
    occ = np.array([2.0, 1.0, 0.5])
    cd, vd = _dev(c), _dev(v)
    res = {}
    for mode in ("static", "generic"):
        if mode == "generic":
            monkeypatch.setenv("QB200_NO_STATIC", "1")
        ft = H.FourierTransform(b, *grid)
        ft.set_workspace(2 << 30)
        assert not ft.fused() and ft.query(14) == 1 and ft.query(15) == (1 if mode == "static" else 0)
        assert ft.query(19) == (2 if mode == "static" else 0)      # static: y stage with two TMEM lanes per 252-point column
        h = torch.zeros_like(cd)
        H.rs_mul_add(ft, cd, vd, h, kpg2=_dev(b["kpg2"]))
        rho = torch.zeros(N, dtype=torch.float64, device="cuda")
        H.compute_density(ft, cd, 1.0, occ, b["omega"], rho)
        res[mode] = (h.cpu().numpy(), rho.cpu().numpy())
        if mode == "static":
            f = torch.zeros(N, dtype=torch.complex128, device="cuda")
            ft.backward(cd[0].contiguous(), f)
            nrm = float((f.abs() ** 2).sum().item()) / N
            assert abs(nrm - float(np.vdot(c[0], c[0]).real)) < 1e-10 * nrm          # Parseval
            back = torch.zeros(ngw, dtype=torch.complex128, device="cuda")
            ft.forward(f, back)
            assert relerr(back.cpu().numpy(), c[0]) < 1e-12                           # round trip
            del f, back
            nel = float(rho.sum().item()) * b["omega"] / N
            want = float(sum(occ[n] * np.vdot(c[n], c[n]).real for n in range(nst)))
            assert abs(nel - want) < 1e-10 * want
        del ft, h, rho
        torch.cuda.empty_cache()
    assert relerr(res["static"][0], res["generic"][0]) < 1e-12
    assert relerr(res["static"][1], res["generic"][1]) < 1e-12


@pytest.mark.parametrize("compiled", [True, "ycols", "smem", False])
def test_cuda_vs_oracle_si54p_shape_gamma_real(compiled, monkeypatch):
    """examples/si54p as a Gamma-point real-wavefunction case (fcc-type cell 2 x 15.525, 65 Ry, 126^3 grid,
    ngw 33114): planes of 126 x 127 x 16 B exceed shared memory, so this runs the real-basis pair path (+ odd tail)
    through the z-column and split xy kernels -- with the compiled 126 x 126 shape and with the run-time-shape engine;
    checked against the oracle on the same seeded inputs."""
    if not compiled:
        monkeypatch.setenv("QB200_NO_STATIC", "1")
    if compiled == "smem":      # the compiled shape with the shared-memory y stage (k_ycols2) instead of the tensor-memory one
        monkeypatch.setenv("QB200_YCOLS_T", "0")
    if compiled == "ycols":     # the split path with the tensor-memory y stage (k_ycols_t) instead of the one-kernel xy stage (k_plane_f)
        monkeypatch.setenv("QB200_PLANE_F", "0")
    a = 15.525
    cell, ecut = (0, a, a, a, 0, a, a, a, 0), 32.5
    b = P.make_basis(cell, ecut, (0, 0, 0), False)
    grid = P.density_grid(cell, ecut)
    assert b["is_real"] and grid == (126, 126, 126) and b["ngw"] == 33114
    nst = 5
    c = R.synth_coefficients(b["kpg2"], ecut, nst, b["ngw"] + 3, True, seed=51)
    v = R.synth_potential(*grid, seed=52)
    occ = np.array([2.0, 2.0, 1.0, 0.0, 0.5])
    oft = P.FT(b, *grid)
    ft = H.FourierTransform(b, *grid)
    assert not ft.fused() and ft.query(14) == 1 and ft.query(11) == 1 and ft.query(15) == (2 if compiled else 0)
    assert ft.query(19) == (1 if compiled in (True, "ycols") else 0), "y stage: k_ycols_t (column in tensor memory) only for the compiled shape"
    assert ft.query(20) == (1 if compiled is True else 0), "xy stage in one kernel (k_plane_f) for the compiled 126 x 126 shape"
    want = oft.rs_mul_add(c, v, np.zeros_like(c))
    P.kinetic_add(b["kpg2"], c, want)
    got = _dev(np.zeros_like(c))
    H.rs_mul_add(ft, _dev(c), _dev(v), got, kpg2=_dev(b["kpg2"]))
    got = got.cpu().numpy()
    assert relerr(got[:, :b["ngw"]], want[:, :b["ngw"]]) < TOL
    assert np.all(got[:, b["ngw"]:] == 0)
    rho_want = oft.compute_density(c, occ / b["omega"], np.zeros(oft.N))
    rho = _dev(np.zeros(oft.N))
    H.compute_density(ft, _dev(c), 1.0, occ, b["omega"], rho)
    assert relerr(rho.cpu().numpy(), rho_want) < TOL


@pytest.mark.parametrize("nst,host,imag_head", [(9, False, False), (8, True, False), (1, False, False), (6, False, True), (6, True, True)])
def test_cuda_density_of_real_states_in_pairs(nst, host, imag_head, monkeypatch):
    """Gamma-point real bases: the density transforms two states at once as psi_1 + i psi_2 and adds fac1 Re^2 + fac2 Im^2
    (the form of SlaterDet.cc:858-885; the odd last state alone, :886-903).  Checked against the oracle's one-state-at-a-time
    sum (SlaterDet.cc:906-924, the branch the reference runs) and against this library with QB200_DENSITY_PAIRS=0, with pairs
    whose first, second or both weights are zero, several batches of pair units, device and host pointers.  imag_head: a state
    whose G = 0 coefficient has an imaginary part (outside the reference's invariant for real bases, SlaterDet.cc:2776-2779)
    would leak it into its partner at first order -- such a block must take the one-state branch and still match the oracle."""
    cell, ecut = (11, 0, 0, 0, 12, 0, 0, 0, 13), 6.0
    b = P.make_basis(cell, ecut, (0, 0, 0), False)
    grid = P.density_grid(cell, ecut)
    assert b["is_real"]
    ldc = b["ngw"] + 2
    c = R.synth_coefficients(b["kpg2"], ecut, nst, ldc, True, seed=71)
    if imag_head:
        c[3, 0] += 1e-3j * abs(c[3, 0])
    occ = np.array([2.0, 0.0, 0.0, 1.5, 0.0, 0.0, 1.0, 2.0, 0.75])[:nst]
    oft = P.FT(b, *grid)
    rho_want = oft.compute_density(c, occ / b["omega"], np.full(oft.N, 0.5))
    res = {}
    for pairs in ("1", "0"):
        monkeypatch.setenv("QB200_DENSITY_PAIRS", pairs)
        ft = H.FourierTransform(b, *grid)
        ft.set_workspace(2 * ft.nvec() * grid[2] * 16)   # two units per batch
        rho = np.full(oft.N, 0.5) if host else _dev(np.full(oft.N, 0.5))
        H.compute_density(ft, c if host else _dev(c), 1.0, occ, b["omega"], rho)
        res[pairs] = rho if host else rho.cpu().numpy()
        assert relerr(res[pairs], rho_want) < TOL
    assert relerr(res["1"], res["0"]) < 1e-12


@pytest.mark.parametrize("kpoint,host", [((0, 0, 0), False), ((0.25, 0, 0.5), True)])
def test_cuda_update_density_tail_vs_oracle(kpoint, host):
    """ChargeDensity::update_density (ChargeDensity.cc:276-551, norm-conserving, one k-point): rho(r), the integral
    nelectrons = sum(rho)*omega/N and rho(G) = vft->forward(omega*rho) on the density basis (k = 0, 4*ecut), against the
    oracle restatement (compute_density + forward transform, both pinned to the reference)."""
    cell, ecut, nst = (10, 0, 0, 0, 11, 0, 0, 0, 12), 5.0, 6
    b = P.make_basis(cell, ecut, kpoint, False)
    vb = P.make_basis(cell, 4.0 * ecut, (0, 0, 0), False)          # ChargeDensity.cc:77-81
    grid = P.density_grid(cell, ecut)
    N = grid[0] * grid[1] * grid[2]
    c = R.synth_coefficients(b["kpg2"], ecut, nst, b["ngw"], b["is_real"], seed=61)
    occ = np.array([2, 2, 2, 1, 0.5, 0.0])
    oft, ovft = P.FT(b, *grid), P.FT(vb, *grid)
    rho_ref = oft.compute_density(c, occ / b["omega"], np.zeros(N))
    nel_ref = rho_ref.sum() * b["omega"] / N
    rhog_ref = ovft.forward((b["omega"] * rho_ref).astype(np.complex128))
    ft = H.FourierTransform(b, *grid)
    cd = H.ChargeDensity(ft, vb, b["omega"])
    if host:
        rhor, rhog, cc = np.zeros(N), np.zeros(vb["ngw"], dtype=np.complex128), c
    else:
        rhor = torch.zeros(N, dtype=torch.float64, device="cuda")
        rhog = torch.zeros(vb["ngw"], dtype=torch.complex128, device="cuda")
        cc = _dev(c)
    nel = cd.update_density(cc, occ, rhor, rhog)
    back = (lambda a: a) if host else (lambda t: t.cpu().numpy())
    assert relerr(back(rhor), rho_ref) < TOL
    assert abs(nel - nel_ref) < 1e-12 * abs(nel_ref)
    assert relerr(back(rhog), rhog_ref) < TOL
    # the G = 0 coefficient is the electron count (vbasis is real: G = 0 sits at index 0, SlaterDet.cc:2776-2779)
    assert abs(back(rhog)[0].real - nel) < 1e-10 * abs(nel)


@pytest.mark.parametrize("host", [False, True])
def test_cuda_exponential_propagator_vs_oracle(host):
    """ExponentialWavefunctionStepper::exponential(2, dt1, dt2) with a frozen Hamiltonian
    (ExponentialWavefunctionStepper.cc:51-149): 4th-order Taylor series of exp(-i dt H) applied to the block, both time
    steps from one chain of H applications; oracle = the same recurrence over the oracle's H psi."""
    cell, ecut, kpoint, nst = (10, 0, 0, 0, 11, 0, 0, 0, 12), 5.0, (0.0, 0.0, 0.0), 5
    b = P.make_basis(cell, ecut, kpoint, True)                       # force_complex_wf ON (vars/WfDyn.h:82-92)
    assert not b["is_real"]
    grid = P.density_grid(cell, ecut)
    ngw = b["ngw"]
    c = R.synth_coefficients(b["kpg2"], ecut, nst, ngw + 1, False, seed=71)
    v = 0.3 * R.synth_potential(*grid, seed=72)
    occ = R.synth_occ(nst, nst)
    rng = np.random.default_rng(73)
    species = [dict(na=2, npr=4, lproj=np.array([0, 1, 1, 1], dtype=np.int32), wt=np.array([0.9, -0.4, -0.4, -0.4]),
                    twnl=rng.standard_normal((4, ngw)) * np.exp(-b["kpg2"] / 4.0)[None, :], tau=rng.uniform(0, 10, (2, 3)))]
    dt1, dt2, order = 0.05, 0.1, 4
    oft = P.FT(b, *grid)
    exp1, exp2, op = c.copy(), c.copy(), c.copy()
    f1 = f2 = 1.0 + 0.0j
    for n in range(1, order + 1):
        f1 *= -1j * dt1 / n
        f2 *= -1j * dt2 / n
        _, op = P.hpsi(b, oft, np.ascontiguousarray(op), v, occ, species)
        exp1 += f1 * op
        exp2 += f2 * op
    ft = H.FourierTransform(b, *grid)
    nlp = H.NonLocalPotential(b, species)
    if host:
        cw, c2w, back = c.copy(), np.zeros_like(c), (lambda a: a)
        H.exponential(ft, nlp, cw, occ, v, b["kpg2"], dt1, dt2, c2w, order)
    else:
        cw, c2w, back = _dev(c), _dev(np.zeros_like(c)), (lambda t: t.cpu().numpy())
        H.exponential(ft, nlp, cw, occ, _dev(v), _dev(b["kpg2"]), dt1, dt2, c2w, order)
    assert relerr(back(cw), exp1) < TOL
    assert relerr(back(c2w), exp2) < TOL
    # norm conservation to the order of the truncated series (the reference's own accuracy check for ETRS)
    n0 = np.vdot(c[0], c[0]).real
    n1 = np.vdot(back(cw)[0], back(cw)[0]).real
    assert abs(n1 - n0) < 1e-4 * n0
    # single-exponential form (num_exp == 1)
    cw1 = _dev(c)
    H.exponential(ft, nlp, cw1, occ, _dev(v), _dev(b["kpg2"]), dt1, order=order)
    assert relerr(cw1.cpu().numpy(), exp1) < TOL


@pytest.mark.parametrize("name", ["forced_complex_ortho_al", "mgo216_shape_112cubed"])
def test_cuda_projectors_gamma_half_sphere_mode(name, monkeypatch):
    """complex states at k = 0 (force_complex_wf): the projector contraction runs over the half sphere as two real
    functions per state (query 14 == 3); the Karatsuba path on the same inputs must agree, and so must H psi accumulated
    into a non-zero block and a second call (cached positions)"""
    g = load_golden(name)
    b = P.make_basis(g["cell"], g["ecut"], tuple(g["kpoint"]), bool(g["force_complex"]))
    c, v, occ = regen_inputs(g, b["kpg2"])
    cd = _dev(c)
    nlp = H.NonLocalPotential(b, g["species"])
    base = np.ascontiguousarray(c[::-1] * 0.5)                  # cp += ...: start from a non-zero block
    cp = _dev(base)
    enl = nlp.energy(cd, occ, True, cp)
    assert nlp.query(14) == 3
    assert abs(enl - g["enl"]) <= 1e-10 * max(1.0, abs(g["enl"]))
    compare(g, "hnl", cp.cpu().numpy() - base)
    cp2 = _dev(base)
    assert nlp.energy(cd, occ, True, cp2) == enl                # deterministic
    assert np.array_equal(cp2.cpu().numpy(), cp.cpu().numpy())
    assert abs(nlp.energy(cd, occ, False) - enl) <= 1e-14 * max(1.0, abs(enl))
    monkeypatch.setenv("QB200_NL_GAMMA", "0")
    nlk = H.NonLocalPotential(b, g["species"])
    cp3 = _dev(base)
    enl3 = nlk.energy(cd, occ, True, cp3)
    assert nlk.query(14) == 2
    assert abs(enl3 - enl) <= 1e-12 * max(1.0, abs(enl))
    assert relerr(cp3.cpu().numpy(), cp.cpu().numpy()) < 1e-12


def test_cuda_projectors_kpoint_keeps_three_product_form():
    g = load_golden("kpoint_cubic_au_oncv")
    b = P.make_basis(g["cell"], g["ecut"], tuple(g["kpoint"]), bool(g["force_complex"]))
    c, v, occ = regen_inputs(g, b["kpg2"])
    nlp = H.NonLocalPotential(b, g["species"])
    nlp.energy(_dev(c), occ, False)
    assert nlp.query(14) == 2


def test_cuda_projectors_asymmetric_tables_use_general_path():
    """the half-sphere form needs twnl(-G) = (-1)^l twnl(G); a table without that property (legal input of the ABI) must be
    detected and go through the general complex path, matching the oracle"""
    cell, ecut = (10, 0, 0, 0, 11, 0, 0, 0, 12), 5.0
    b = P.make_basis(cell, ecut, (0, 0, 0), True)
    ngw = b["ngw"]
    c = R.synth_coefficients(b["kpg2"], ecut, 6, ngw, False, seed=5)
    occ = R.synth_occ(6, 5)
    rng = np.random.default_rng(9)
    species = [dict(na=2, npr=4, lproj=np.array([0, 1, 1, 1], dtype=np.int32), wt=np.array([0.9, -0.4, -0.4, -0.4]),
                    twnl=rng.standard_normal((4, ngw)) * np.exp(-b["kpg2"] / 4.0)[None, :], tau=rng.uniform(0, 10, (2, 3)))]
    enl_ref, h_ref = P.nl_energy(b, c, occ, species)
    nlp = H.NonLocalPotential(b, species)
    cp = _dev(np.zeros_like(c))
    enl = nlp.energy(_dev(c), occ, True, cp)
    assert nlp.query(14) == 2
    assert abs(enl - enl_ref) <= 1e-10 * max(1.0, abs(enl_ref)) and relerr(cp.cpu().numpy(), h_ref) < TOL


def test_cuda_energy_after_fixed_tddft_steps_within_1e8_hartree():
    """north_star: energies within 1e-8 Ha of the reference after a fixed TDDFT step count.  The terms this path owns --
    E_kin + E_nl + integral(v rho) = sum_n occ_n <psi_n|H|psi_n>, plus E_nl alone and the electron count -- after 5
    propagation steps (order-4 exponential, Hamiltonian frozen, real MgO-type projector tables from the fixture so that the
    Gamma half-sphere projector form is the one exercised) against the same recurrence over the oracle."""
    g = load_golden("forced_complex_ortho_al")
    b = P.make_basis(g["cell"], g["ecut"], tuple(g["kpoint"]), True)
    c, v, occ = regen_inputs(g, b["kpg2"])
    grid = (g["np0"], g["np1"], g["np2"])
    species, ngw = g["species"], b["ngw"]
    v = 0.3 * v
    dt, order, nsteps = 0.05, 4, 5
    oft = P.FT(b, *grid)
    cref = c.copy()
    for _ in range(nsteps):
        acc, op, f = cref.copy(), cref.copy(), 1.0 + 0.0j
        for n in range(1, order + 1):
            f *= -1j * dt / n
            _, op = P.hpsi(b, oft, np.ascontiguousarray(op), v, occ, species)
            acc += f * op
        cref = acc
    enl_ref, h_ref = P.hpsi(b, oft, np.ascontiguousarray(cref), v, occ, species)
    e_ref = float(np.sum(occ * np.einsum("ng,ng->n", cref[:, :ngw].conj(), h_ref[:, :ngw]).real))
    rho_ref = oft.compute_density(np.ascontiguousarray(cref), occ / b["omega"], np.zeros(oft.N))
    ft = H.FourierTransform(b, *grid)
    nlp = H.NonLocalPotential(b, species)
    cd, vd, kd = _dev(c), _dev(v), _dev(b["kpg2"])
    for _ in range(nsteps):
        H.exponential(ft, nlp, cd, occ, vd, kd, dt, order=order)
    assert nlp.query(14) == 3
    out = torch.zeros_like(cd)
    enl = H.hpsi(ft, nlp, cd, occ, vd, kd, out)
    cg, hg = cd.cpu().numpy(), out.cpu().numpy()
    e = float(np.sum(occ * np.einsum("ng,ng->n", cg[:, :ngw].conj(), hg[:, :ngw]).real))
    rho = torch.zeros(oft.N, dtype=torch.float64, device="cuda")
    H.compute_density(ft, cd, 1.0, occ, b["omega"], rho)
    nel, nel_ref = float(rho.sum()) * b["omega"] / oft.N, float(rho_ref.sum()) * b["omega"] / oft.N
    assert abs(e - e_ref) < 1e-8 and abs(enl - enl_ref) < 1e-8 and abs(nel - nel_ref) < 1e-8, (e - e_ref, enl - enl_ref, nel - nel_ref)
    assert relerr(cg, cref) < TOL


def test_cuda_gamma_half_sphere_host_blocks_sliced():
    """the half-sphere projector form under the host-pointer pipeline: 259 complex states at k = 0 go through in slices of 64
    (the anl matrix of the first slice is reused by the later ones); tables with the parity of real projectors
    (l = 0: even in G, l = 1: odd); against the device-pointer path (one block) and the oracle"""
    cell, ecut, nst = (11, 0, 0, 0, 12, 0, 0, 0, 13), 6.0, 259
    b = P.make_basis(cell, ecut, (0, 0, 0), True)
    grid = P.density_grid(cell, ecut)
    ngw = b["ngw"]
    c = R.synth_coefficients(b["kpg2"], ecut, nst, ngw + 2, False, seed=61)
    v = R.synth_potential(*grid, seed=62)
    occ = R.synth_occ(nst, nst - 7)
    env = np.exp(-b["kpg2"] / 4.0)
    twnl = np.stack([env, env * b["kpgx"][0], env * b["kpgx"][1], env * b["kpgx"][2]])
    rng = np.random.default_rng(63)
    species = [dict(na=5, npr=4, lproj=np.array([0, 1, 1, 1], dtype=np.int32), wt=np.array([1.3, -0.6, -0.6, -0.6]),
                    twnl=np.ascontiguousarray(twnl), tau=rng.uniform(0, 11, (5, 3)))]
    ft = H.FourierTransform(b, *grid)
    nlp = H.NonLocalPotential(b, species)
    out_d = _dev(np.zeros_like(c))
    enl_d = H.hpsi(ft, nlp, _dev(c), occ, _dev(v), _dev(b["kpg2"]), out_d)
    assert nlp.query(14) == 3
    hc = torch.from_numpy(c.copy()).pin_memory()
    hout = torch.zeros_like(hc).pin_memory()
    enl_h = H.hpsi(ft, nlp, hc, occ, v, b["kpg2"], hout)
    assert nlp.query(14) == 3
    assert abs(enl_h - enl_d) <= 1e-12 * max(1.0, abs(enl_d))
    assert relerr(hout.numpy(), out_d.cpu().numpy()) < 1e-12
    assert np.all(hout.numpy()[:, ngw:] == 0)
    sel = [0, 1, 130, nst - 2, nst - 1]
    oft = P.FT(b, *grid)
    enl_ref, _ = P.nl_energy(b, c, occ, species, compute_hpsi=False)
    _, h_ref = P.hpsi(b, oft, np.ascontiguousarray(c[sel]), v, occ[sel], species)
    assert abs(enl_h - enl_ref) <= 1e-10 * max(1.0, abs(enl_ref))
    assert relerr(hout.numpy()[sel][:, :ngw], h_ref[:, :ngw]) < TOL


# ------------------------------------------------------------------------------------------------ the benchmark's own regime
@pytest.mark.parametrize("mode", ["half_sphere", "three_product", "host_pointers"])
def test_cuda_mgo216_all_atoms_benchmark_projector_regime(mode, monkeypatch):
    """examples/MgO216 as the benchmark runs it: every atom of mg108o108h1.sys with the shipped Mg/O/H potentials --
    540 projector rows in two species (Mg 108 x 4 at row offset 0, O 108 x 1 at offset 432 > 128, H local only), i.e.
    five 128-row M tiles with a padded last tile -- 16 complex states on the 112^3 grid, against the arrays the reference
    itself produced (NonLocalPotential.cc:1909-2171 with its atom blocks of :1541; fixture mgo216_all_atoms_16st).
    Half-sphere form (k_split_pm / k_fnl<1> / k_back<2>, W streamed once), the general 3-product form, and host pointers."""
    if mode == "three_product":
        monkeypatch.setenv("QB200_NL_GAMMA", "0")
    _run_fixture("mgo216_all_atoms_16st", mode != "host_pointers", False, monkeypatch)
    g = load_golden("mgo216_all_atoms_16st")
    assert [s["na"] * s["npr"] for s in g["species"]] == [432, 108, 0]
    b = P.make_basis(g["cell"], g["ecut"], tuple(g["kpoint"]), bool(g["force_complex"]))
    c, v, occ = regen_inputs(g, b["kpg2"])
    nlp = H.NonLocalPotential(b, g["species"])
    # a block wider than one 64-state GEMM tile: the fixture's 16 states repeated with different occupations; column n of
    # V_nl psi must not depend on its neighbours and E_nl is additive over states
    reps = 5
    cw = _dev(np.concatenate([c] * reps))
    occw = np.concatenate([occ * (1.0 + 0.25 * r) for r in range(reps)])
    cp = torch.zeros_like(cw)
    enl = nlp.energy(cw, occw, True, cp)
    assert nlp.query(13) == 540 and nlp.query(14) == (2 if mode == "three_product" else 3)
    want = g["enl"] * sum(1.0 + 0.25 * r for r in range(reps))
    assert abs(enl - want) <= 1e-10 * abs(want)
    got = cp.cpu().numpy()
    for r in range(reps):
        compare(g, "hnl", got[16 * r:16 * (r + 1)])


def _many_atoms_species(b, rng, shapes):
    """projector tables with the parity of real projectors (twnl(-G) = (-1)^l twnl(G)) for many atoms"""
    kpg = np.sqrt(b["kpg2"])
    out = []
    for na, lproj in shapes:
        rows = []
        for i, l in enumerate(lproj):
            ang = 1.0 if l == 0 else (b["kpgx"][i % 3] / np.maximum(kpg, 1e-12)) ** l
            rows.append((1.0 + 0.07 * i) * np.exp(-b["kpg2"] / (3.0 + 0.2 * i)) * ang)
        a = np.array(b["cell"], dtype=np.float64).reshape(3, 3)
        out.append(dict(na=na, npr=len(lproj), lproj=np.array(lproj, dtype=np.int32),
                        wt=rng.uniform(0.4, 1.6, len(lproj)) * np.where(np.array(lproj) == 1, -1.0, 1.0),
                        twnl=np.ascontiguousarray(np.array(rows)), tau=rng.uniform(0, 1, (na, 3)) @ a))
    return out


@pytest.mark.parametrize("kpoint,fc", [((1e-7, 0.0, 0.0), False), ((0.0, 0.0, 0.0), True), ((0.0, 0.0, 0.0), False)])
def test_cuda_projectors_many_rows_chunked_vs_oracle(kpoint, fc):
    """the gold benchmark's projector regime in miniature: 710 projector rows (two species, row offset 360 of the second,
    a multi-channel l <= 2 species), a workspace that forces >= 3 plane-wave chunks per sweep (the anl block is regenerated
    in both sweeps), 40 states; complex basis at a tiny k (Au992's k = (1e-7,0,0), 3-product form), complex states at
    Gamma (half-sphere form) and real states at Gamma, against the oracle (NonLocalPotential.cc:1909-2171)"""
    cell, ecut, nst = (14, 0, 0, 0, 15, 0, 0, 0, 21), 7.0, 40
    b = P.make_basis(cell, ecut, kpoint, fc)
    b["cell"] = cell
    ngw = b["ngw"]
    rng = np.random.default_rng(91)
    species = _many_atoms_species(b, rng, [(90, [0, 1, 1, 1]), (35, [0, 0, 1, 1, 1, 2, 2, 2, 2, 2])])
    c = R.synth_coefficients(b["kpg2"], ecut, nst, ngw + 1, b["is_real"], seed=92)
    occ = R.synth_occ(nst, nst - 3)
    enl_ref, h_ref = P.nl_energy(b, c, occ, species)
    nlp = H.NonLocalPotential(b, species)
    nlp.set_workspace(710 * 24 * (ngw // 4))            # about a quarter of the sphere per chunk
    base = np.ascontiguousarray(0.25 * c[::-1])
    cp = _dev(base)
    enl = nlp.energy(_dev(c), occ, True, cp)
    assert nlp.query(13) == 710 and nlp.query(11) >= 3, (nlp.query(13), nlp.query(11))
    assert abs(enl - enl_ref) <= 1e-10 * max(1.0, abs(enl_ref)), (enl, enl_ref)
    assert relerr(cp.cpu().numpy() - base, h_ref) < TOL
    # unchunked on the same inputs: identical E_nl to rounding, and H psi through the whole-Hpsi entry point
    nl1 = H.NonLocalPotential(b, species)
    cp1 = _dev(base)
    enl1 = nl1.energy(_dev(c), occ, True, cp1)
    assert nl1.query(11) == 1 and abs(enl1 - enl) <= 1e-12 * max(1.0, abs(enl))
    assert relerr(cp1.cpu().numpy(), cp.cpu().numpy()) < 1e-12


def test_cuda_bulkal_fcc_kpoint_properties():
    """examples/bulkal (bulkal_kp3.i): fcc primitive cell, k = (0, 0, 1/3), 20 Ry, 20^3 grid, Al with l <= 3; the arrays are
    pinned by the reference fixture bulkal_fcc_kpoint (parametrised tests above); here the shape the survey records
    (ngw 153, 37 rods) and H = H^dagger on the block"""
    g = load_golden("bulkal_fcc_kpoint")
    assert (g["np0"], g["np1"], g["np2"]) == (20, 20, 20) and g["ngw"] == 153 and g["nrods"] == 37 and not g["is_real"]
    assert sorted(set(int(l) for l in g["species"][0]["lproj"])) == [0, 1, 3]      # lmax 3, llocal 2
    b = P.make_basis(g["cell"], g["ecut"], tuple(g["kpoint"]), False)
    c, v, occ = regen_inputs(g, b["kpg2"])
    ft = H.FourierTransform(b, 20, 20, 20)
    nlp = H.NonLocalPotential(b, g["species"])
    cd = _dev(c)
    out = torch.zeros_like(cd)
    H.hpsi(ft, nlp, cd, occ, _dev(v), _dev(b["kpg2"]), out)
    m = (cd[:, :g["ngw"]].conj() @ out[:, :g["ngw"]].T).cpu().numpy()
    assert np.abs(m - m.conj().T).max() < 1e-11 * np.abs(m).max()


@pytest.mark.parametrize("kpoint,fc,host", [((0, 0, 0), False, False), ((0.1, 0.2, 0.3), False, True), ((0, 0, 0), True, False)])
def test_cuda_ekin_sums_vs_oracle(kpoint, fc, host):
    """kinetic-energy section of EnergyFunctional::energy (EnergyFunctional.cc:1155-1296): psi2sum and the 14 partial sums
    (ekin, sigma_ekin, econf, sigma_econf) against the oracle's loop-for-loop restatement; trace(sigma_ekin sums) = 2 ekin"""
    cell, ecut, nst = (10, 0, 0, 0.5, 11, 0, 0, 0, 12), 6.0, 37
    b = P.make_basis(cell, ecut, kpoint, fc)
    grid = P.density_grid(cell, ecut)
    ngw = b["ngw"]
    c = R.synth_coefficients(b["kpg2"], ecut, nst, ngw + 3, b["is_real"], seed=81)
    occ = R.synth_occ(nst, nst - 4)
    rng = np.random.default_rng(82)
    fstress, dfstress = rng.uniform(0, 1, ngw), rng.uniform(-1, 1, ngw)
    ft = H.FourierTransform(b, *grid)
    w = (lambda a: np.ascontiguousarray(a)) if host else _dev
    t_ref, p2_ref = P.ekin_sums(b["kpg2"], c, occ, b["is_real"], kpgx=b["kpgx"], fstress=fstress, dfstress=dfstress)
    t, p2 = H.ekin_sums(ft, w(c), occ, b["is_real"], w(b["kpg2"]), w(b["kpgx"]), w(fstress), w(dfstress), want_psi2sum=True)
    assert relerr(p2, p2_ref) < 1e-13
    assert np.abs(t - t_ref).max() < 1e-12 * np.abs(t_ref).max(), (t, t_ref)
    assert abs(t[1] + t[2] + t[3] - 2.0 * t[0]) < 1e-12 * abs(t[0])
    # energy only: no stress / confinement tables
    t0 = H.ekin_sums(ft, w(c), occ, b["is_real"], w(b["kpg2"]))
    assert abs(t0[0] - t_ref[0]) < 1e-12 * abs(t_ref[0]) and np.all(t0[1:] == 0.0)
    # E_kin of the block equals the kinetic part of sum_n occ_n <psi_n|H psi_n> (real basis: G and -G)
    kin = np.zeros_like(c)
    P.kinetic_add(b["kpg2"], c, kin)
    e = np.einsum("ng,ng->n", c[:, :ngw].conj(), kin[:, :ngw]).real
    if b["is_real"]:
        e = 2.0 * e - (c[:, 0].conj() * kin[:, 0]).real
    assert abs(t0[0] - float(np.dot(occ, e))) < 1e-12 * abs(t0[0])


def test_cuda_hpsi_without_host_synchronisation_keeps_enl_on_the_device():
    """qb200_hpsi with enl = NULL does not synchronise; qb200_nl_last_enl hands the energy out afterwards -- to a device address
    (behind the density, so that one all-reduce carries rho and E_nl) or to the host; same value as the synchronous call"""
    g = load_golden("kpoint_cubic_au_oncv")
    b = P.make_basis(g["cell"], g["ecut"], tuple(g["kpoint"]), bool(g["force_complex"]))
    c, v, occ = regen_inputs(g, b["kpg2"])
    ft = H.FourierTransform(b, g["np0"], g["np1"], g["np2"])
    nlp = H.NonLocalPotential(b, g["species"])
    cd, vd, kd = _dev(c), _dev(v), _dev(b["kpg2"])
    out1, out2 = _dev(np.zeros_like(c)), _dev(np.zeros_like(c))
    e_sync = H.hpsi(ft, nlp, cd, occ, vd, kd, out1)
    assert H.hpsi(ft, nlp, cd, occ, vd, kd, out2, want_enl=False) is None
    N = g["np0"] * g["np1"] * g["np2"]
    ext = torch.zeros(N + 8, dtype=torch.float64, device="cuda")
    nlp.last_enl(ext[N:N + 1])
    assert nlp.last_enl() == e_sync and float(ext[N].item()) == e_sync and abs(e_sync - g["enl"]) <= 1e-10 * max(1.0, abs(g["enl"]))
    assert torch.equal(out1, out2)


def test_cuda_plans_of_different_shared_memory_need_coexist():
    """regression (found by running the reference itself through the shim, examples/sih4): the wavefunction plan and the
    density-basis plan of one run share the plane kernels but need different amounts of dynamic shared memory; the
    opt-in is per kernel, so the plan created LATER (smaller need) must not undercut launches of the earlier one"""
    cell, ecut = (14, 0, 0, 0, 14, 0, 0, 0, 14), 18.0
    b = P.make_basis(cell, ecut, (0, 0, 0), False)
    vb = P.make_basis(cell, 4.0 * ecut, (0, 0, 0), False)
    grid = P.density_grid(cell, ecut)
    assert grid == (60, 60, 60)
    N = 60 ** 3
    vft = H.FourierTransform(vb, *grid)          # larger column table in shared memory
    ft = H.FourierTransform(b, *grid)            # created second, needs less
    rng = np.random.default_rng(5)
    f = rng.standard_normal(N) + 0j
    want = P.FT(vb, *grid).forward(f.copy())
    got = np.zeros(vb["ngw"], dtype=np.complex128)
    vft.forward(f.copy(), got)
    assert relerr(got, want) < TOL
    c = R.synth_coefficients(b["kpg2"], ecut, 2, b["ngw"], True, seed=3)
    fr = np.zeros(N, dtype=np.complex128)
    ft.backward(np.ascontiguousarray(c[0, :b["ngw"]]), fr)
    assert relerr(fr, P.FT(b, *grid).backward(np.ascontiguousarray(c[0, :b["ngw"]]))) < TOL
