"""CPU: the package's host-side Basis mirror (qball_b200/basis.py) is bit-identical to the oracle's (which is pinned
to the reference's Basis by the golden fixtures), and the C-ABI library loads and exports every declared symbol."""
import ctypes
import os
import re

import numpy as np
import pytest

import port as P
from qball_b200 import basis as B
from util import golden_names, load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("cell,ecut,kp,fc", [
    ((14, 0, 0, 0, 14, 0, 0, 0, 14), 18.0, (0, 0, 0), False),
    ((23.1, 0, 0, 0, 23.1, 0, 0, 0, 23.1), 25.0, (0, 0, 0), True),
    ((7.2, 0, 0, 0.6, 8.1, 0, 0.3, -0.5, 9.3), 4.5, (0, 0, 0), False),
    ((7.7, 0, 0, 0, 7.7, 0, 0, 0, 7.7), 4.0, (0.25, 0.1, -0.3), False),
    ((5.4, 5.4, 0, 0, 5.4, 5.4, 5.4, 0, 5.4), 8.0, (0.125, 0.25, 0.375), False),
])
def test_host_basis_matches_oracle_bitwise(cell, ecut, kp, fc):
    a, b = B.make_basis(cell, ecut, kp, fc), P.make_basis(cell, ecut, kp, fc)
    for k in ("is_real", "ngw", "nrods", "idxmin1", "idxmax1", "omega"):
        assert a[k] == b[k], k
    assert tuple(a["basis_np"]) == tuple(b["basis_np"])
    for k in ("rod_h", "rod_k", "rod_lmin", "rod_size", "idx", "kpg2", "kpgx"):
        assert np.array_equal(a[k], b[k]), k
    assert B.density_grid(cell, ecut) == P.density_grid(cell, ecut)


@pytest.mark.parametrize("name", golden_names())
def test_host_basis_matches_reference_fixture(name):
    g = load_golden(name)
    a = B.make_basis(g["cell"], g["ecut"], tuple(g["kpoint"]), bool(g["force_complex"]))
    for k in ("rod_h", "rod_k", "rod_lmin", "rod_size"):
        assert np.array_equal(a[k], g[k]), k
    assert a["omega"] == g["omega"] and a["ngw"] == g["ngw"]
    if g["mode"] == "full":
        assert np.array_equal(a["kpg2"], g["kpg2"]) and np.array_equal(a["kpgx"], g["kpgx"])


def test_c_abi_library_exports_every_declared_symbol():
    from qball_b200 import build, capi
    build.build()
    hdr = open(os.path.join(ROOT, "include", "qball_b200.h")).read()
    names = sorted(set(re.findall(r"\b(qb200_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 20
    lib = ctypes.CDLL(capi.lib_path())
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/qball_b200.h but not exported"
    # no compute without a GPU: the product must fail loudly, not fall back
    L = capi.load()
    if L.qb200_device_count() == 0:
        h = ctypes.c_void_p()
        z = (ctypes.c_int * 1)(0)
        s = (ctypes.c_int * 1)(3)
        rc = L.qb200_plan_create(ctypes.byref(h), 0, 8, 8, 8, 1, z, z, z, s, 0, 0, 0)
        assert rc < 0 and L.qb200_last_error()


def test_device_side_synthetic_block_is_bit_identical_to_the_numpy_generator():
    """bench.py synthesises the coefficient block directly in HBM (torch int64 splitmix); it must be the same block the
    CPU arm, the fixtures and the parity leg generate with numpy"""
    from qball_b200 import basis as B
    from qball_b200 import synth as S
    b = B.make_basis((9, 0, 0, 0, 10, 0, 0, 0, 11), 5.0, (0, 0, 0), False)
    for is_real, first in ((b["is_real"], 0), (False, 7)):
        a = S.synth_coefficients(b["kpg2"], 5.0, 4, b["ngw"] + 2, is_real, seed=3, first_state=first)
        t = S.synth_coefficients_torch(b["kpg2"], 5.0, 4, b["ngw"] + 2, is_real, seed=3, first_state=first).numpy()
        assert np.array_equal(a, t)
