"""CPU: the plain-C oracle against the compiled reference run LIVE (oracle/_ref/ref_driver), where that binary exists
(the build container; it also travels to the GPU box).  Covers shapes the committed fixtures do not."""
import numpy as np
import pytest

import port as P
import refdrive as R
from util import TOL, relerr

pytestmark = pytest.mark.skipif(not R.have_ref(), reason="oracle/_ref/ref_driver not built (needs /root/reference)")


def _species_from_fixture(name):
    from util import load_golden
    return load_golden(name)


@pytest.mark.parametrize("cell,ecut,kpoint,fc,nst", [
    ((10, 0, 0, 0, 10, 0, 0, 0, 10), 6.0, (0, 0, 0), False, 3),
    ((10, 0, 0, 0, 10, 0, 0, 0, 10), 6.0, (0, 0, 0), True, 2),
    ((8, 0, 0, 2.0, 9, 0, -1.0, 0.5, 12), 5.0, (0.5, 0.5, 0.5), False, 2),
    ((5.4, 5.4, 0, 0, 5.4, 5.4, 5.4, 0, 5.4), 8.0, (0.125, 0.25, 0.375), False, 1),
])
def test_port_vs_live_reference_no_species(cell, ecut, kpoint, fc, nst):
    case = R.Case(cell=cell, ecut=ecut, kpoint=kpoint, force_complex=fc, nst=nst)
    r = R.run_reference(case, seed=3)
    b = P.make_basis(cell, ecut, kpoint, fc)
    for k in ("rod_h", "rod_k", "rod_lmin", "rod_size", "idx", "kpg2", "kpgx"):
        assert np.array_equal(b[k], r[k]), k
    assert P.density_grid(cell, ecut) == (r["np0"], r["np1"], r["np2"])
    ft = P.FT(b, r["np0"], r["np1"], r["np2"])
    c, v, occ = r["c"], r["v"], r["occ"]
    assert relerr(ft.backward(c[0, :b["ngw"]]), r["bwd0"]) < TOL
    assert relerr(ft.rs_mul_add(c, v, np.zeros_like(c)), r["hloc"]) < TOL
    assert relerr(ft.compute_density(c, occ / r["omega"], np.zeros(ft.N)), r["rho"]) < TOL
    enl, h = P.hpsi(b, ft, c, v, occ, [])
    assert relerr(h, r["hpsi"]) < TOL
