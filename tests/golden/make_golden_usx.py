"""tests/golden/make_golden_usx.py -- golden vectors for the rest of the ultrasoft path: the ultrasoft
branch of NonLocalPotential::energy (NonLocalPotential.cc:1554-1752: D_nm^I from veff(G) and Q_nm(G), E_nl, H psi) and the
augmentation charges of ChargeDensity::update_density (ChargeDensity.cc:312-465), produced by the reference's own classes on a
Sample (oracle/_ref/ref_driver mode `usx`) with the ultrasoft carbon potential of the reference's test suite, seeded
coefficients, fractional occupations and a seeded veff(G).  Stored: the tables the reference produced (betag, Q_nm(G) on the
density basis, dzero, channels, positions -- inputs of the C ABI), veff, and the results (betapsi, enl, hnl, rho without and with
the augmentation charges).
    make -C oracle ref && python tests/golden/make_golden_usx.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import refdrive as R  # noqa: E402

CARBON = "/root/reference/testsuite/pseudopotentials/04_ultrasoft_carbon/carbon.xml"
CASES = {
    # two carbon atoms off the grid points, Gamma point (ultrasoft forces complex states), one empty state
    "carbon2_gamma": (R.Case(cell=(8, 0, 0, 0, 8.5, 0, 0, 0, 9), ecut=4.0, force_complex=True, nst=3, species=[("carbon", CARBON)],
                             atoms=[("C1", "carbon", 0.3, 0.2, -0.1), ("C2", "carbon", 2.4, -1.1, 0.7)]), 3, [2.0, 1.0, 0.0]),
    # three atoms, a k-point, a triclinic cell, fractional occupations
    "carbon3_kpoint": (R.Case(cell=(8, 0, 0, 1.0, 8.5, 0, 0.5, -0.7, 9), ecut=4.5, kpoint=(0.25, 0.0, -0.125), nst=4,
                              species=[("carbon", CARBON)],
                              atoms=[("C1", "carbon", 0.0, 0.0, 0.0), ("C2", "carbon", 2.4, -1.1, 0.7), ("C3", "carbon", -3.0, 3.0, 1.2)]),
                       7, [2.0, 2.0, 0.5, 1.5]),
}


def main():
    outdir = os.path.join(HERE, "usx")
    os.makedirs(outdir, exist_ok=True)
    for name, (case, seed, occ) in CASES.items():
        r = R.run_reference_usx(case, seed=seed, occ=occ)
        d = dict(cell=np.array(case.cell, dtype=np.float64), ecut=case.ecut, kpoint=np.array(case.kpoint), nst=case.nst, seed=seed,
                 occ=r["occ"], ngw=r["ngw"], mloc=r["mloc"], omega=r["omega"], nsp=len(r["species"]), vgrid=np.array(r["vgrid"]),
                 ngv=r["ngv"], veff=r["veff"], enl=r["enl"], hnl=r["hnl"], rho_nc=r["rho_nc"], rho=r["rho"], nelectrons=r["nelectrons"])
        for i, s in enumerate(r["species"]):
            for k in ("na", "npr", "nq", "lproj", "twnl", "tau", "lm1", "lm2", "dzero", "qnmg", "betapsi"):
                d[f"sp{i}_{k}"] = s[k]
        fn = os.path.join(outdir, name + ".npz")
        np.savez_compressed(fn, **d)
        print(f"{name}: ngw {r['ngw']} ngv {r['ngv']} grid {r['vgrid']} nst {case.nst} -> {os.path.getsize(fn) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
