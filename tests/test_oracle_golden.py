"""CPU: the plain-C oracle (oracle/qb_oracle.c) against the committed golden vectors produced by the reference itself."""
import numpy as np
import pytest

import port as P
from util import compare, golden_names, load_golden, regen_inputs


def _basis_from_golden(g):
    b = P.make_basis(g["cell"], g["ecut"], tuple(g["kpoint"]), bool(g["force_complex"]))
    # basis tables must be IDENTICAL to the reference's (integer work: bit exact)
    for k in ("ngw", "nrods", "is_real", "idxmin1", "idxmax1"):
        assert b[k] == g[k], k
    assert tuple(b["basis_np"]) == tuple(int(x) for x in g["basis_np"])
    for k in ("rod_h", "rod_k", "rod_lmin", "rod_size"):
        assert np.array_equal(b[k], g[k]), k
    assert b["omega"] == g["omega"]
    return b


@pytest.mark.parametrize("name", golden_names())
def test_port_matches_reference_fixture(name):
    g = load_golden(name)
    b = _basis_from_golden(g)
    if g["mode"] == "full":
        assert np.array_equal(b["kpg2"], g["kpg2"]) and np.array_equal(b["kpgx"], g["kpgx"])
    c, v, occ = regen_inputs(g, b["kpg2"])
    ft = P.FT(b, g["np0"], g["np1"], g["np2"])
    ngw = g["ngw"]
    f = ft.backward(c[0, :ngw])
    compare(g, "bwd0", f)
    compare(g, "fwd0", ft.forward(f * v))
    if g["is_real"] and g["nst"] >= 2:
        f = ft.backward(c[0, :ngw], c[1, :ngw])
        compare(g, "bwdpair01", f)
        c1, c2 = ft.forward(f * v, pair=True)
        compare(g, "fwdpair0", c1)
        compare(g, "fwdpair1", c2)
    compare(g, "hloc", ft.rs_mul_add(c, v, np.zeros_like(c)))
    compare(g, "rho", ft.compute_density(c, occ / g["omega"], np.zeros(ft.N)))
    enl, hnl = P.nl_energy(b, c, occ, g["species"])
    assert abs(enl - g["enl"]) <= 1e-10 * max(1.0, abs(g["enl"]))
    if g["nsp"]:
        compare(g, "hnl", hnl)
    enl2, h = P.hpsi(b, ft, c, v, occ, g["species"])
    compare(g, "hpsi", h)


def test_density_grid_rule():
    # ChargeDensity.cc:88-99 -- sizes confirmed by running the reference
    assert P.density_grid((14, 0, 0, 0, 14, 0, 0, 0, 14), 18.0) == (60, 60, 60)
    assert P.density_grid((23.1, 0, 0, 0, 23.1, 0, 0, 0, 23.1), 25.0) == (112, 112, 112)


def test_factorizable():
    L = P.lib()
    ok = [n for n in range(2, 200) if L.qbo_factorizable(n)]
    assert 112 in ok and 126 in ok and 60 in ok and 154 in ok  # 154 = 2*7*11
    assert 26 not in ok and 34 not in ok and 50 not in ok and 98 not in ok  # 13, 17, 5^2, 7^2


def test_transform_identities():
    """KATs lifted from src/tests/testFourierTransform.cc:297-365: fwd(bwd(c)) = c; a single plane wave has |f|=1."""
    b = P.make_basis((9, 0, 0, 0, 9, 0, 0, 0, 9), 5.0, (0.3, 0.0, 0.0), False)
    g = P.density_grid((9, 0, 0, 0, 9, 0, 0, 0, 9), 5.0)
    ft = P.FT(b, *g)
    rng = np.random.default_rng(0)
    c = rng.standard_normal(b["ngw"]) + 1j * rng.standard_normal(b["ngw"])
    assert np.abs(ft.forward(ft.backward(c)) - c).max() < 1e-13
    e = np.zeros(b["ngw"], dtype=complex)
    e[7] = 1.0
    assert np.abs(np.abs(ft.backward(e)) - 1.0).max() < 1e-13
