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


def main():
    outdir = os.path.join(HERE, "twnl")
    os.makedirs(outdir, exist_ok=True)
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
