"""Neighbour list: CUDA (vssr_nbr_build) vs oracle — bit-exact (rowptr, col, shift)."""
import numpy as np
import pytest
import torch

from conftest import perturbed, with_adsorbates
from oracle.nbrlist import neighbor_list, to_csr

pytestmark = pytest.mark.gpu


def _oracle_batch(structs, rc):
    rowptrs, cols, shifts, base, ebase = [np.zeros(1, np.int32)], [], [], 0, 0
    for s in structs:
        i, j, S = neighbor_list(s["positions"], s["cell"], s["pbc"], rc)
        rp = to_csr(i, len(s["numbers"]))
        rowptrs.append(rp[1:] + ebase)
        cols.append(j + base)
        shifts.append(S)
        base += len(s["numbers"])
        ebase += len(i)
    return np.concatenate(rowptrs), np.concatenate(cols), np.concatenate(shifts)


def _gpu(structs, rc, e_cap=None):
    from surface_sampling_b200 import engine
    b = engine.Batch.from_arrays([s["positions"] for s in structs], [s["numbers"] for s in structs],
                                 [s["cell"] for s in structs], [s["pbc"] for s in structs])
    rowptr, col, shift = engine.neighbor_list(b, rc, e_cap=e_cap)
    torch.cuda.synchronize()
    return rowptr.cpu().numpy(), col.cpu().numpy(), shift.cpu().numpy()[:, :3].astype(np.int32)


@pytest.mark.parametrize("rc", [3.1, 5.0, 6.0])
def test_all_fixture_slabs_bit_exact(structures, rc):
    structs = [structures[n] for n in sorted(structures)]
    rp, col, sh = _gpu(structs, rc)
    rp0, col0, sh0 = _oracle_batch(structs, rc)
    assert np.array_equal(rp, rp0) and np.array_equal(col, col0) and np.array_equal(sh, sh0)


def test_perturbed_adsorbates_unwrapped_and_capacity_retry(structures):
    rng = np.random.default_rng(3)
    base = structures["SrTiO3_001_2x2"]
    structs = []
    for k in range(6):
        s = with_adsorbates(perturbed(base, rng, 0.08), rng, k, [8, 22, 38])
        if k % 2:  # push some atoms outside the unit cell (unwrapped positions)
            s["positions"] = s["positions"] + np.array([9.0, -17.0, 0.0]) * (rng.random((len(s["numbers"]), 1)) > 0.7)
        structs.append(s)
    structs.append(structures["GaN_0001_3x3"])   # hexagonal cell, pbc TTF
    structs.append(structures["SrTiO3_unit_cell"])  # 3.9 A cell: many images per pair
    rp, col, sh = _gpu(structs, 6.0, e_cap=1000)  # too small on purpose: engine must retry
    rp0, col0, sh0 = _oracle_batch(structs, 6.0)
    assert np.array_equal(rp, rp0) and np.array_equal(col, col0) and np.array_equal(sh, sh0)


def test_empty_and_single_atom_structures(structures):
    s1 = {"positions": np.zeros((1, 3)), "numbers": np.array([8]), "cell": np.eye(3) * 4.0, "pbc": np.array([True] * 3)}
    s0 = {"positions": np.zeros((0, 3)), "numbers": np.zeros(0, int), "cell": np.eye(3) * 4.0, "pbc": np.array([True] * 3)}
    structs = [s0, s1, structures["Au_110_2x2"], s0]
    rp, col, sh = _gpu(structs, 5.0)
    rp0, col0, sh0 = _oracle_batch(structs, 5.0)
    assert np.array_equal(rp, rp0) and np.array_equal(col, col0) and np.array_equal(sh, sh0)
    assert rp[1] - rp[0] == 6  # the lone atom sees its 6 nearest periodic images (4 A < 5 A < 4*sqrt(2) A)
