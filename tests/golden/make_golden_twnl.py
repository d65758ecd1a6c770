"""tests/golden/make_golden_twnl.py -- what NonLocalPotential::update_twnl builds twnl FROM, dumped
from the reference's own Species objects by oracle/_ref/ref_driver for the Kleinman-Bylander species of the golden cases of
make_golden.py: per projector its m and radial-table index, and the species' radial cubic splines (knots, values, second
derivatives: Species::gspl_, Spline::y_, y2_ of Species::projectors_g_[l][ic]) truncated to the knots that cover the case's
max |k+G| (+ the last knot of the full table, beyond which Species::dvnlg returns 0).  The twnl tables themselves are in
the main fixtures (tests/golden/<case>.npz), produced by the reference's update_twnl.
    make -C oracle ref && python tests/golden/make_golden_twnl.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
sys.path.insert(0, HERE)
import refdrive as R  # noqa: E402
import make_golden as MG  # noqa: E402

NAMES = ["gamma_triclinic_si_h", "kpoint_cubic_au_oncv", "forced_complex_ortho_al", "bulkal_fcc_kpoint", "mgo216_shape_112cubed"]


PP = "/root/reference/share/pseudopotentials/quantum-simulation.org/hscv/pbe/"
# semi-local species (nquad = 8): Yb (lmax 3, llocal 3: l = 0, 1, 2 projectors, 72 in all), Zr (lmax 2, llocal 0: l = 1, 2; 64);
# a k-point in a triclinic cell and the Gamma point (the q = 0 row of the l = 0 Bessel function).  These cases are not in the
# main fixtures, so the reference's tables are stored here with the basis parameters.
SEMILOCAL = {
    "semilocal_yb_kpoint": R.Case(cell=(8.2, 0, 0, 0.4, 7.9, 0, 0.2, -0.3, 8.8), ecut=4.0, kpoint=(0.1, 0.0, -0.2), nst=2, species=[("ytterbium", PP + "Yb_HSCV_PBE-1.0.xml")],
                                  atoms=[("Yb1", "ytterbium", 0.3, 0.2, -0.1), ("Yb2", "ytterbium", 3.9, 4.1, 3.7)]),
    "semilocal_yb_gamma": R.Case(cell=(8.2, 0, 0, 0, 7.9, 0, 0, 0, 8.8), ecut=4.0, nst=2, species=[("ytterbium", PP + "Yb_HSCV_PBE-1.0.xml")],
                                 atoms=[("Yb1", "ytterbium", 0.3, 0.2, -0.1)]),
    "semilocal_zr_kpoint": R.Case(cell=(8.2, 0, 0, 0.4, 7.9, 0, 0.2, -0.3, 8.8), ecut=4.0, kpoint=(0.1, 0.0, -0.2), nst=2, species=[("zirconium", PP + "Zr_HSCV_PBE-1.1.xml")],
                                  atoms=[("Zr1", "zirconium", 0.3, 0.2, -0.1)]),
}


def main():
    outdir = os.path.join(HERE, "twnl")
    os.makedirs(outdir, exist_ok=True)
    sl = os.path.join(HERE, "twnl_semilocal")
    os.makedirs(sl, exist_ok=True)
    for name, case in SEMILOCAL.items():
        r = R.run_reference(case, seed=5)
        s = r["species"][0]
        assert s["nquad"] > 0
        fn = os.path.join(sl, name + ".npz")
        np.savez_compressed(fn, cell=np.array(case.cell, dtype=np.float64), ecut=case.ecut, kpoint=np.array(case.kpoint), ngw=r["ngw"], omega=r["omega"],
                            na=s["na"], npr=s["npr"], nquad=s["nquad"], lproj=s["lproj"], wt=s["wt"], tau=s["tau"], mproj=s["mproj"], rproj=s["rproj"],
                            twnl=s["twnl"], seed=5, nst=case.nst, mloc=r["mloc"], enl=r["enl"], hnl_checksum=float(np.abs(r["hnl"]).sum()))
        print(f"{name}: npr {s['npr']} ngw {r['ngw']} -> {os.path.getsize(fn) / 1024:.0f} KiB")
    for name in NAMES:
        case, seed, nocc, mode, stride = MG.CASES[name]
        r = R.run_reference(case, seed=seed, nocc=nocc)
        kmax = float(np.sqrt(r["kpg2"].max()))       # kpg2 = |k+G|^2 (Basis.cc:733)
        d = dict(nsp=len(r["species"]))
        for i, s in enumerate(r["species"]):
            d[f"sp{i}_kb"] = int(s["npr"] > 0 and s["nquad"] == 0)
            if not d[f"sp{i}_kb"]:
                continue
            K = int(np.searchsorted(s["gspl"], kmax * 1.02)) + 2
            d.update({f"sp{i}_mproj": s["mproj"], f"sp{i}_tabproj": s["tabproj"], f"sp{i}_gspl": s["gspl"][:K], f"sp{i}_gcut": s["gspl"][-1],
                      f"sp{i}_vnlg": s["vnlg"][:, :K], f"sp{i}_vnlg_spl": s["vnlg_spl"][:, :K]})
            if mode == "sampled":       # the main fixture holds only a sample of twnl: keep a checksum and a strided sample of the table
                d[f"sp{i}_twnl_sum"] = float(np.abs(s["twnl"]).sum())
                d[f"sp{i}_twnl_stride"] = 97
                d[f"sp{i}_twnl_sample"] = s["twnl"][:, ::97].copy()
        fn = os.path.join(outdir, name + ".npz")
        np.savez_compressed(fn, **d)
        print(f"{name}: {os.path.getsize(fn) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
