"""tests/golden/make_golden_us.py -- golden vectors for the ultrasoft beta.psi path, produced by the
reference's own SlaterDet::init_usfns (calc_betag, calc_betapsi, Species::calc_qnmg -> set_qaug, calc_spsi; SlaterDet.cc:103-197,
2006-2263, 2426-2570) in oracle/_ref/ref_driver (mode `us`) on the ultrasoft carbon potential of the reference's own test
testsuite/pseudopotentials/04_ultrasoft_carbon, seeded coefficients (oracle/refdrive.py).  Stored: the betag tables the
reference produced (inputs of the C ABI, like twnl), l per channel, the (lm1, lm2, qaug) triples, positions, betapsi, spsi.
    make -C oracle ref && python tests/golden/make_golden_us.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import refdrive as R  # noqa: E402

CARBON = "/root/reference/testsuite/pseudopotentials/04_ultrasoft_carbon/carbon.xml"
CASES = {
    # two carbon atoms off the grid points in a cubic cell, Gamma point (ultrasoft forces complex states)
    "carbon2_gamma": (R.Case(cell=(12, 0, 0, 0, 12, 0, 0, 0, 12), ecut=12.0, force_complex=True, nst=3, species=[("carbon", CARBON)],
                             atoms=[("C1", "carbon", 0.3, 0.2, -0.1), ("C2", "carbon", 2.4, -1.1, 0.7)]), 3),
    # three atoms, a k-point, a triclinic cell
    "carbon3_kpoint": (R.Case(cell=(11, 0, 0, 1.5, 10, 0, 0.5, -1.0, 12), ecut=10.0, kpoint=(0.25, 0.0, -0.125), nst=4,
                              species=[("carbon", CARBON)],
                              atoms=[("C1", "carbon", 0.0, 0.0, 0.0), ("C2", "carbon", 2.4, -1.1, 0.7), ("C3", "carbon", -3.0, 4.0, 1.2)]), 7),
}


def main():
    outdir = os.path.join(HERE, "us")
    os.makedirs(outdir, exist_ok=True)
    for name, (case, seed) in CASES.items():
        r = R.run_reference_us(case, seed=seed)
        d = dict(cell=np.array(case.cell, dtype=np.float64), ecut=case.ecut, kpoint=np.array(case.kpoint), force_complex=int(case.force_complex),
                 nst=case.nst, seed=seed, ngw=r["ngw"], mloc=r["mloc"], omega=r["omega"], nsp=len(r["species"]), spsi=r["spsi"])
        for i, s in enumerate(r["species"]):
            for k in ("na", "npr", "lproj", "twnl", "tau", "lm1", "lm2", "qaug", "betapsi"):
                d[f"sp{i}_{k}"] = s[k]
        fn = os.path.join(outdir, name + ".npz")
        np.savez_compressed(fn, **d)
        print(f"{name}: ngw {r['ngw']} nst {case.nst} -> {os.path.getsize(fn) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
