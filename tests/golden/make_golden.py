"""tests/golden/make_golden.py -- regenerates the committed golden fixtures from the reference itself.

Runs ONLY where /root/reference and oracle/_ref/ref_driver exist (the build container):
    make -C oracle ref && python tests/golden/make_golden.py
Each fixture holds the reference's basis tables, projector tables (twnl/wt/lproj/tau are outputs of the reference's
host-side setup, NonLocalPotential.cc:76-1522) and the arrays produced by the reference's own FourierTransform,
SlaterDet::rs_mul_add, SlaterDet::compute_density and NonLocalPotential::energy on the deterministic synthetic inputs of
oracle/refdrive.py (regenerated from the stored seed by the tests; a checksum guards against generator drift).
"full" fixtures store whole arrays (small grids); "sampled" fixtures store strided samples + sums (benchmark-size grids).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import refdrive as R  # noqa: E402

REF = "/root/reference"
EX = REF + "/examples"
SI = EX + "/sih4/Si_PBE.xml"
H = EX + "/sih4/H_PBE.xml"
AL = EX + "/bulkal/TM_CAPW91_Al.xml"
AU_ONCV = REF + "/testsuite/pseudopotentials/05_gold_oncv/Au_ONCV_PBE-1.0.xml"
MG = EX + "/MgO216/Mg.xml"
OX = EX + "/MgO216/O.xml"
HY = EX + "/MgO216/H.xml"


def sys_atoms(path):
    """atom lines of a reference .sys file -> [(name, species, x, y, z)]"""
    out = []
    for line in open(path):
        w = line.split()
        if w and w[0] == "atom":
            out.append((w[1], w[2], float(w[3]), float(w[4]), float(w[5])))
    return out


CASES = {
    # name: (Case, seed, nocc, mode, stride)
    "gamma_triclinic_si_h": (R.Case(cell=(7.2, 0, 0, 0.6, 8.1, 0, 0.3, -0.5, 9.3), ecut=4.5, nst=3,
                                    species=[("silicon", SI), ("hydrogen", H)],
                                    atoms=[("Si1", "silicon", 0.1, 0.2, -0.1), ("Si2", "silicon", 3.3, 4.1, 4.4),
                                           ("H1", "hydrogen", 1.6, 1.6, 1.6)]), 11, 2, "full", 1),
    "kpoint_cubic_au_oncv": (R.Case(cell=(7.7, 0, 0, 0, 7.7, 0, 0, 0, 7.7), ecut=4.0, nst=2, kpoint=(0.25, 0.1, -0.3),
                                    species=[("gold", AU_ONCV)],
                                    atoms=[("Au1", "gold", 0.0, 0.0, 0.0), ("Au2", "gold", 3.85, 3.85, 3.7)]), 12, None,
                             "full", 1),
    "forced_complex_ortho_al": (R.Case(cell=(6.4, 0, 0, 0, 7.6, 0, 0, 0, 9.9), ecut=4.0, nst=2, force_complex=True,
                                       species=[("aluminum", AL)],
                                       atoms=[("Al1", "aluminum", 0.3, 0.1, 0.2), ("Al2", "aluminum", 3.2, 3.8, 5.0),
                                              ("Al3", "aluminum", -1.0, 2.0, -3.0)]), 13, None, "full", 1),
    "gamma_even_owngrid_si": (R.Case(cell=(7.0, 0, 0, 0, 7.0, 0, 0, 0, 8.0), ecut=4.0, nst=4, grid=(24, 20, 30),
                                     species=[("silicon", SI)], atoms=[("Si1", "silicon", 1.1, -0.2, 0.7)]), 14, 3, "full", 1),
    # benchmark-size grids: sih4 as shipped, and the MgO216 basis/grid with a handful of atoms
    "sih4_60cubed": (R.Case(cell=(14, 0, 0, 0, 14, 0, 0, 0, 14), ecut=18.0, nst=5,
                            species=[("silicon", SI), ("hydrogen", H)],
                            atoms=[("Si1", "silicon", 0.0, 0.0, 0.0), ("H1", "hydrogen", 1.6, 1.6, 1.6),
                                   ("H2", "hydrogen", -1.6, -1.6, 1.6), ("H3", "hydrogen", 1.6, -1.6, -1.6),
                                   ("H4", "hydrogen", -1.6, 1.6, -1.6)]), 15, 4, "sampled", 61),
    "mgo216_shape_112cubed": (R.Case(cell=(23.1, 0, 0, 0, 23.1, 0, 0, 0, 23.1), ecut=25.0, nst=2, force_complex=True,
                                     species=[("magnesium", MG), ("oxygen", OX)],
                                     atoms=[("Mg1", "magnesium", 0.0, 0.0, 0.0), ("Mg2", "magnesium", 3.85, 3.85, 0.0),
                                            ("Mg3", "magnesium", 11.55, 7.7, 3.85), ("O1", "oxygen", 3.85, 0.0, 0.0),
                                            ("O2", "oxygen", 0.0, 3.85, 0.0)]), 16, None, "sampled", 997),
    # the benchmark's own projector regime: every atom of examples/MgO216/mg108o108h1.sys with the
    # shipped Mg/O/H potentials (Mg 108 x 4 + O 108 x 1 = 540 projector rows in two species, H local only), 16 states
    "mgo216_all_atoms_16st": (R.Case(cell=(23.1, 0, 0, 0, 23.1, 0, 0, 0, 23.1), ecut=25.0, nst=16, force_complex=True,
                                     species=[("magnesium", MG), ("oxygen", OX), ("hydrogen", HY)],
                                     atoms=sys_atoms(EX + "/MgO216/mg108o108h1.sys")), 17, 14, "sampled", 2039),
    # examples/bulkal/bulkal_kp3.i: fcc primitive cell, second k-point (0, 0, 1/3), 20 Ry, l <= 3 (lmax 3, llocal 2)
    "bulkal_fcc_kpoint": (R.Case(cell=(-3.8, 0, 3.8, 0, 3.8, 3.8, -3.8, 3.8, 0), ecut=10.0, nst=6, kpoint=(0.0, 0.0, 0.3333333),
                                 species=[("aluminum", AL)], atoms=[("Al1", "aluminum", 0.0, 0.0, 0.0)]), 18, 4, "full", 1),
}


def checksum(a: np.ndarray) -> float:
    a = np.ascontiguousarray(a).view(np.float64).ravel()
    w = np.cos(np.arange(a.size, dtype=np.float64) * 0.7853981633974483 + 0.3)
    return float(np.dot(a, w))


def main():
    outdir = os.path.dirname(os.path.abspath(__file__))
    only = sys.argv[1:]          # fixture names to (re)generate; default: all
    for name, (case, seed, nocc, mode, stride) in CASES.items():
        if only and name not in only:
            continue
        r = R.run_reference(case, seed=seed, nocc=nocc)
        d = dict(mode=mode, stride=stride, seed=seed, nocc=-1 if nocc is None else nocc,
                 cell=np.array(case.cell, dtype=np.float64), ecut=case.ecut, kpoint=np.array(case.kpoint, dtype=np.float64),
                 force_complex=int(case.force_complex), grid=np.array([r["np0"], r["np1"], r["np2"]]), nst=case.nst,
                 mloc=r["mloc"], ngw=r["ngw"], nrods=r["nrods"], is_real=int(r["is_real"]),
                 basis_np=np.array(r["basis_np"]), idxmin1=r["idxmin1"], idxmax1=r["idxmax1"],
                 rod_h=r["rod_h"], rod_k=r["rod_k"], rod_lmin=r["rod_lmin"], rod_size=r["rod_size"], omega=r["omega"],
                 nsp=r["nsp"], enl=r["enl"],
                 in_checksum=np.array([checksum(r["c"]), checksum(r["v"]), checksum(r["occ"])]))
        full = mode == "full"
        if full:
            d["kpg2"] = r["kpg2"]
            d["kpgx"] = r["kpgx"]
        else:
            d["kpg2_checksum"] = checksum(r["kpg2"])
            d["kpgx_checksum"] = checksum(r["kpgx"])
        for i, s in enumerate(r["species"]):
            d[f"sp{i}_na"], d[f"sp{i}_npr"] = s["na"], s["npr"]
            d[f"sp{i}_lproj"], d[f"sp{i}_wt"], d[f"sp{i}_tau"] = s["lproj"], s["wt"], s["tau"]
            tw = s["twnl"].astype(np.float64)
            if full or tw.size < 70000:
                d[f"sp{i}_twnl"] = tw
            else:
                # large tables: a cubic cell repeats every value many times (twnl = Y_lm(G) v_l(|G|)); store each row's
                # distinct values + the index of every plane wave into them (bit-exact, ~20x smaller)
                u, inv = np.unique(tw, return_inverse=True)
                d[f"sp{i}_twnl_uniq"] = u
                d[f"sp{i}_twnl_inv"] = inv.reshape(tw.shape).astype(np.uint32)
        keys = ["bwd0", "fwd0", "hloc", "rho", "hpsi"] + (["hnl"] if r["nsp"] else [])
        if r["is_real"] and case.nst >= 2:
            keys += ["bwdpair01", "fwdpair0", "fwdpair1"]
        for k in keys:
            a = r[k]
            if full:
                d[k] = a
            else:
                flat = a.ravel()
                d[k + "_sample"] = flat[::stride].copy()
                d[k + "_sum"] = np.array([flat.sum(), np.abs(flat).max(), np.vdot(flat, flat).real])
        fn = os.path.join(outdir, name + ".npz")
        np.savez_compressed(fn, **d)
        print(f"{name}: grid {r['np0']}x{r['np1']}x{r['np2']} ngw {r['ngw']} nvec-rods {r['nrods']} enl {r['enl']:.12f} "
              f"-> {os.path.getsize(fn) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
