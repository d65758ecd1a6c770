"""tests/golden/make_golden_la.py -- golden vectors for the subspace dense linear algebra,
produced by the reference's own matrix classes in oracle/_ref/ref_driver:
  resid, resid_a : DoubleMatrix/ComplexMatrix gemm (+ ger) exactly as PSDAWavefunctionStepper::update calls them
                   (PSDAWavefunctionStepper.cc:65-84, 264-277) on (psi, H psi) of the fixture
  gram           : SlaterDet::gram() (SlaterDet.cc:1043-1143) on psi
  diag_w         : eigenvalues of h = psi^H (H psi) by the reference's own syevd / heev('l') as Wavefunction::diag calls them
                   (Wavefunction.cc:1538-1539, 1612, 1641, 1693)
  cur            : SlaterDet::compute_density(ft, w, complex* rho, sd2) (SlaterDet.cc:935-968) driven as
                   CurrentDensity::update_current does (CurrentDensity.cc:52-86), weight 1
Same cases, seeds and synthetic inputs as make_golden.py ("full" fixtures), stored under tests/golden/la/.
    make -C oracle ref && python tests/golden/make_golden_la.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import CASES, R, checksum  # noqa: E402


def main():
    outdir = os.path.join(HERE, "la")
    os.makedirs(outdir, exist_ok=True)
    only = sys.argv[1:]          # fixture names to (re)generate; default: all "full" cases
    for name, (case, seed, nocc, mode, stride) in CASES.items():
        if mode != "full" or (only and name not in only):
            continue
        r = R.run_reference(case, seed=seed, nocc=nocc)
        d = dict(resid=r["resid"], resid_a=r["resid_a"], gram=r["gram"], cur=r["cur"], hpsi_checksum=checksum(r["hpsi"]),
                 diag_w=r["diag_w"])
        fn = os.path.join(outdir, name + ".npz")
        np.savez_compressed(fn, **d)
        print(f"{name}: nst {case.nst} ngw {r['ngw']} -> {os.path.getsize(fn) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
