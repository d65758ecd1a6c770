"""The reference's `save -fast` checkpoint (Wavefunction::write_fast / read_fast, Wavefunction.cc:2296-2403, 2651-2788) as the
on-disk exchange format of the per-GPU state blocks (the format part).  The fixture was WRITTEN BY THE
REFERENCE (tests/golden/make_golden_fast.py: examples/sih4 at 12 Ry, LDA, 4 + 2 states, converged); the tests read it with
qball_b200.fastio, check it against what the reference printed for that wavefunction, and write it back byte for byte."""
import json
import os

import numpy as np
import pytest

import port as P
from qball_b200 import fastio
from util import GOLDEN

FIX = os.path.join(GOLDEN, "fast", "sih4_lda_12ry")
HARTREE_EV = 27.2114            # the conversion the reference prints eigenvalues with


def _load():
    meta = json.load(open(FIX + ".json"))
    blocks = fastio.read_fast(fastio.task_file(FIX, 0), meta["mloc"], meta["nst"], has_empty=True)
    assert len(blocks) == 1
    return meta, blocks[0]


def test_read_reference_checkpoint_and_write_it_back(tmp_path):
    meta, blk = _load()
    c = blk["c"]
    assert c.shape == (meta["nst"], meta["mloc"]) and meta["converged"]
    b = P.make_basis(tuple(meta["cell"]), meta["ecut_hartree"], (0, 0, 0), False)
    assert b["ngw"] == meta["ngw"] and b["is_real"]
    # what the reference knows about these states: occupations, eigenvalues (printed in eV), orthonormality after gram
    assert np.array_equal(blk["occ"], [2, 2, 2, 2, 0, 0])
    assert np.abs(blk["eig"] * HARTREE_EV - np.array(meta["eigenvalues_ev"])).max() < 2e-4
    s = P.subspace_h(c, c, True)                                      # real-basis overlap: 2 c^T c - row-0 term
    assert np.abs(s - np.eye(meta["nst"])).max() < 1e-12
    # E_kin as EnergyFunctional::energy sums it (EnergyFunctional.cc:1155-1296) = the <ekin> the reference printed
    ekin = P.ekin_sums(b["kpg2"], np.ascontiguousarray(c), blk["occ"], True)[0][0]
    assert abs(ekin - meta["ekin"]) < 5e-8, (ekin, meta["ekin"])
    out = tmp_path / "copy000000"
    fastio.write_fast(str(out), [blk])
    assert open(out, "rb").read() == open(fastio.task_file(FIX, 0), "rb").read()
    with pytest.raises(ValueError):
        fastio.read_fast(fastio.task_file(FIX, 0), meta["mloc"], meta["nst"], has_empty=False)      # trailing eig / occ


@pytest.mark.gpu
def test_cuda_on_a_reference_checkpoint():
    """the checkpoint block goes into the C ABI as is: E_kin (qb200_ekin_sums) and the electron count of the density
    (qb200_compute_density) reproduce what the reference printed for it (<ekin>, total_electronic_charge: 8.00000000)"""
    import torch
    from qball_b200 import host as H
    meta, blk = _load()
    cell = tuple(meta["cell"])
    b = P.make_basis(cell, meta["ecut_hartree"], (0, 0, 0), False)
    grid = P.density_grid(cell, meta["ecut_hartree"])
    ft = H.FourierTransform(b, *grid)
    cd = torch.from_numpy(np.ascontiguousarray(blk["c"])).cuda()
    ekin = H.ekin_sums(ft, cd, blk["occ"], True, b["kpg2"])[0]
    assert abs(ekin - meta["ekin"]) < 5e-8
    N = grid[0] * grid[1] * grid[2]
    rho = torch.zeros(N, dtype=torch.float64, device="cuda")
    H.compute_density(ft, cd, 1.0, blk["occ"], b["omega"], rho)
    nel = float(rho.sum()) * b["omega"] / N
    assert abs(nel - meta["total_electronic_charge"]) < 1e-8
