"""tests/golden/make_golden_xc.py -- golden vectors of the reference's own exchange-correlation functionals
(LDAFunctional::xc_unpolarized, functionals/LDAFunctional.cc:96-161; PBEFunctional::excpbe, functionals/PBEFunctional.cc:196-291)
produced by oracle/_ref/ref_driver `xc` on a fixed set of (rho, grad rho) points that spans both branches of the
Perdew-Zunger fit (rs < 1, rs >= 1), the rho < 1e-18 cut-off of PBE and reduced gradients s from 0 to ~50.
    make -C oracle ref && python tests/golden/make_golden_xc.py
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import refdrive as R  # noqa: E402


def points(n=1500, seed=7):
    u = R.splitmix_uniform(seed, 0, 5 * n).reshape(5, n)
    rho = 10.0 ** (-20.0 + 22.5 * u[0])                 # 1e-20 .. 3e2
    rho[:8] = [0.0, 1e-19, 1e-18, 0.99e-18, 0.2387324146, 0.2387324147, 1.0, 1e-30]   # cut-offs and the rs = 1 seam
    kf = (3.0 * np.pi ** 2 * rho) ** (1.0 / 3.0)
    s = 10.0 ** (-4.0 + 5.7 * u[1])                     # reduced gradient 1e-4 .. 50
    s[::7] = 0.0
    g = 2.0 * kf * rho * s
    d = u[2:5] - 0.5
    d /= np.maximum(np.sqrt((d ** 2).sum(axis=0)), 1e-30)
    return rho, d * g


def run_reference(rho, grad):
    tmp = tempfile.mkdtemp(prefix="qbxc_")
    pre = os.path.join(tmp, "xc")
    rho.tofile(pre + ".in_rho.f64")
    np.ascontiguousarray(grad).tofile(pre + ".in_grad.f64")
    subprocess.run([R.REF_DRIVER, "xc", pre], check=True)
    return {k: np.fromfile(f"{pre}.{k}.f64") for k in ("lda_exc", "lda_vxc", "pbe_exc", "pbe_vxc1", "pbe_vxc2")}


def main():
    rho, grad = points()
    out = run_reference(rho, grad)
    fn = os.path.join(HERE, "xc", "xc_points.npz")
    np.savez_compressed(fn, rho=rho, grad=grad, **out)
    print(f"xc_points: {rho.size} points -> {os.path.getsize(fn) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
