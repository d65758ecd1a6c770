"""tests/golden/make_golden_fast.py -- a `save -fast` checkpoint written by the reference itself (oracle/_ref/qball):
examples/sih4 at 12 Ry, LDA, 4 occupied + 2 empty states, converged (Harris-Foulkes energy constant to 1e-11 over 10 SCF
steps), together with what the reference printed for that wavefunction (E_kin, electron count, eigenvalues).
    make -C oracle ref && python tests/golden/make_golden_fast.py
(The serial oracle build of the reference crashes AFTER the wavefunction file is complete, while writing its `.lastrhor`
side file; the exit status is therefore ignored and the file size checked instead.)"""
import json
import os
import re
import shutil
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/examples/sih4"
INPUT = """sih4.sys
set ecut 12.0
set wf_dyn PSDA
set xc LDA
set ecutprec 4.0
set nempty 2
set threshold_scf 1.E-11 10
randomize_wf
run 0 80 3
save -fast wf_sih4
quit
"""


def main():
    tmp = tempfile.mkdtemp(prefix="qbfast_")
    for f in ("sih4.sys", "Si_PBE.xml", "H_PBE.xml"):
        shutil.copy(os.path.join(REF, f), tmp)
    open(os.path.join(tmp, "in.i"), "w").write(INPUT)
    r = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "qball"), "in.i"], cwd=tmp, capture_output=True, text=True,
                       env=dict(os.environ, OMP_NUM_THREADS="8"))
    out = r.stdout
    ngw = int(re.search(r"basis size: (\d+)", out).group(1))
    mloc, nst = (int(x) for x in re.search(r"c dimensions: (\d+)x(\d+)", out).groups())
    src = os.path.join(tmp, "wf_sih4000000")
    assert os.path.getsize(src) == 16 * mloc * nst + 16 * nst
    dst = os.path.join(HERE, "fast", "sih4_lda_12ry000000")
    shutil.copy(src, dst)
    eig_ev = [float(x) for x in re.findall(r"<eigenvalues[^>]*>\s*([-0-9.\s]+)</eigenvalues>", out)[-1].split()]
    meta = dict(cell=[14, 0, 0, 0, 14, 0, 0, 0, 14], ecut_hartree=6.0, ngw=ngw, mloc=mloc, nst=nst, nempty=2,
                ekin=float(re.findall(r"<ekin>\s*([-0-9.]+)", out)[-1]), etotal=float(re.findall(r"<etotal>\s*([-0-9.]+)", out)[-1]),
                total_electronic_charge=float(re.findall(r"total_electronic_charge: ([0-9.]+)", out)[-1]), eigenvalues_ev=eig_ev,
                converged="scf convergence" in out)
    json.dump(meta, open(os.path.join(HERE, "fast", "sih4_lda_12ry.json"), "w"), indent=1)
    print(meta)
    shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
