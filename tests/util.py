"""Shared helpers for the parity tests: load a golden fixture and regenerate its synthetic inputs."""
import glob
import os

import numpy as np

import refdrive as R  # oracle/refdrive.py (test infrastructure)

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-10  # north_star: 1e-10 relative on H psi and rho(r)


def golden_names(mode=None):
    out = []
    for f in sorted(glob.glob(os.path.join(GOLDEN, "*.npz"))):
        if mode is None or str(np.load(f)["mode"]) == mode:
            out.append(os.path.basename(f)[:-4])
    return out


def checksum(a):
    a = np.ascontiguousarray(a).view(np.float64).ravel()
    w = np.cos(np.arange(a.size, dtype=np.float64) * 0.7853981633974483 + 0.3)
    return float(np.dot(a, w))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    g = {k: z[k] for k in z.files}
    for k in ("stride", "seed", "nocc", "force_complex", "nst", "mloc", "ngw", "nrods", "is_real", "idxmin1", "idxmax1", "nsp"):
        g[k] = int(g[k])
    for k in ("ecut", "omega", "enl"):
        g[k] = float(g[k])
    g["mode"] = str(g["mode"])
    g["np0"], g["np1"], g["np2"] = (int(x) for x in g["grid"])
    g["is_real"] = bool(g["is_real"])
    def twnl(i):
        # large tables are stored as distinct values + per-plane-wave index (tests/golden/make_golden.py): bit-exact
        if f"sp{i}_twnl" in g:
            return g[f"sp{i}_twnl"]
        return g[f"sp{i}_twnl_uniq"][g[f"sp{i}_twnl_inv"]]

    g["species"] = [dict(na=int(g[f"sp{i}_na"]), npr=int(g[f"sp{i}_npr"]), lproj=g[f"sp{i}_lproj"], wt=g[f"sp{i}_wt"],
                         twnl=twnl(i).reshape(int(g[f"sp{i}_npr"]), g["ngw"]), tau=g[f"sp{i}_tau"].reshape(-1, 3))
                    for i in range(g["nsp"])]
    return g


def regen_inputs(g, kpg2):
    """the deterministic synthetic inputs the fixture was generated from (checked against the stored checksums)"""
    c = R.synth_coefficients(kpg2, g["ecut"], g["nst"], g["mloc"], g["is_real"], g["seed"])
    v = R.synth_potential(g["np0"], g["np1"], g["np2"], g["seed"] + 6)
    occ = R.synth_occ(g["nst"], None if g["nocc"] < 0 else g["nocc"])
    # libm/SIMD exp and cos may differ in the last ulp between hosts: compare the checksums relative to the L1 mass
    for a, want in zip((c, v, occ), g["in_checksum"]):
        l1 = float(np.abs(np.ascontiguousarray(a).view(np.float64)).sum())
        assert abs(checksum(a) - float(want)) <= 1e-13 * max(l1, 1.0), "synthetic input generator drifted"
    return c, v, occ


def relerr(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    den = np.abs(b).max()
    return float(np.abs(a - b).max() / (den if den > 0 else 1.0))


def compare(g, key, arr, tol=TOL):
    """compare a computed array with the fixture (full array, or strided sample + sums)"""
    flat = np.asarray(arr).ravel()
    if g["mode"] == "full":
        ref = g[key].ravel()
        assert flat.shape == ref.shape, (key, flat.shape, ref.shape)
        e = relerr(flat, ref)
    else:
        ref = g[key + "_sample"]
        scale = float(np.abs(g[key + "_sum"][1]))
        e = float(np.abs(flat[::g["stride"]] - ref).max() / scale)
        s = g[key + "_sum"]
        e = max(e, abs(flat.sum() - s[0]) / (abs(s[1]) * np.sqrt(flat.size)), abs(np.vdot(flat, flat).real - s[2].real) / abs(s[2].real))
    assert e < tol, f"{key}: rel err {e:.3e} >= {tol}"
    return e
