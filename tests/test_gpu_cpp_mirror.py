"""include/qball_b200.hpp (the C++ host mirror a Qball maintainer links against): compiles on the CPU box; on the GPU
it runs with plain host std::complex arrays and must match the oracle."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

import port as P
import refdrive as R
from qball_b200 import build, capi
from util import TOL, relerr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _compile(out):
    build.build()
    lib = capi.lib_path()
    cmd = ["g++", "-std=c++11", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "hpp_driver.cc"),
           "-o", out, lib, f"-Wl,-rpath,{os.path.dirname(lib)}", "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64", "-lcudart"]
    subprocess.run(cmd, check=True)


def test_cpp_mirror_compiles_and_links():
    with tempfile.TemporaryDirectory() as d:
        _compile(os.path.join(d, "drv"))


@pytest.mark.gpu
def test_cpp_mirror_matches_oracle():
    cell, ecut, nst = (10, 0, 0, 0, 11, 0, 0, 0, 9), 5.0, 3
    b = P.make_basis(cell, ecut)
    grid = P.density_grid(cell, ecut)
    ldc = b["ngw"] + 2
    c = R.synth_coefficients(b["kpg2"], ecut, nst, ldc, b["is_real"], seed=8)
    v = R.synth_potential(*grid, seed=4)
    occ = R.synth_occ(nst, 2)
    with tempfile.TemporaryDirectory() as d:
        exe, fin, fout = os.path.join(d, "drv"), os.path.join(d, "in.bin"), os.path.join(d, "out.bin")
        _compile(exe)
        with open(fin, "wb") as f:
            np.array([*grid, b["nrods"], int(b["is_real"]), b["idxmin1"], b["idxmax1"], b["ngw"], ldc, nst], dtype=np.int32).tofile(f)
            for k in ("rod_h", "rod_k", "rod_lmin", "rod_size"):
                b[k].astype(np.int32).tofile(f)
            c.tofile(f); v.tofile(f); b["kpg2"].tofile(f); occ.tofile(f); np.array([b["omega"]]).tofile(f)
        subprocess.run([exe, fin, fout], check=True)
        N = grid[0] * grid[1] * grid[2]
        raw = np.fromfile(fout, dtype=np.float64)
        fr = raw[:2 * N].view(np.complex128)
        cp = raw[2 * N:2 * N + 2 * ldc * nst].view(np.complex128).reshape(nst, ldc)
        rho = raw[2 * N + 2 * ldc * nst:3 * N + 2 * ldc * nst]
        tail = raw[3 * N + 2 * ldc * nst:].view(np.complex128).reshape(2, nst, ldc)
    oft = P.FT(b, *grid)
    assert relerr(fr, oft.backward(c[0, :b["ngw"]])) < TOL
    want = oft.rs_mul_add(c, v, np.zeros_like(c))
    P.kinetic_add(b["kpg2"], c, want)
    assert relerr(cp, want) < TOL
    assert relerr(rho, oft.compute_density(c, occ / b["omega"], np.zeros(N))) < TOL
    res_ref, _ = P.residual(c, want, b["is_real"])
    assert relerr(tail[0], res_ref) < TOL
    assert relerr(tail[1], P.gram(c, b["is_real"])) < TOL
