"""ORACLE (test infrastructure only — never imported by the product path).

CPU restatements (fp64, PyTorch autograd for forces) of the classical many-body potentials
the reference evaluates through LAMMPS (un-vendored conda dependency, unpinned:
environment.yml:5-8).  Reference call sites: ``LAMMMPSCalc.run_lammps_calc``
(mcmc/calculators/calculators.py:507-598) with the templates
``tutorials/data/GaN_0001/GaN_0001_lammps_*_template.txt`` (``pair_style tersoff``) and
``tutorials/data/Si_111_5x5/*`` (``pair_style kim`` Stillinger-Weber family).

* Tersoff: LAMMPS ``pair_style tersoff``.  PINNED: pristine GaN(0001) slab
  -144.059 eV (tutorials/GaN_0001.ipynb:228), see tests/test_oracle_golden.py.
* Stillinger-Weber: LAMMPS ``pair_style sw`` functional form with a parameter struct.
  PARITY UNPINNED: the KIM model parameters are not in the reference tree and the Si notebook
  is a missing blob; SW-1985 literature values are used.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch

from .nbrlist import neighbor_list

TERSOFF_ORDER = ["m", "gamma", "lambda3", "c", "d", "costheta0", "n", "beta", "lambda2",
                 "B", "R", "D", "lambda1", "A"]


class TersoffParams:
    """elem3param table: params[(e1,e2,e3)] -> 14 numbers in file order."""

    def __init__(self, pot_json: dict, elements: list[str]):
        self.elements = list(elements)  # type index -> symbol
        ne = len(elements)
        self.table = np.zeros((ne, ne, ne, 14))
        found = np.zeros((ne, ne, ne), dtype=bool)
        for e in pot_json["entries"]:
            try:
                a, b, c = (self.elements.index(x) for x in e["elements"])
            except ValueError:
                continue
            self.table[a, b, c] = e["params"]
            found[a, b, c] = True
        assert found.all(), "missing Tersoff entries"
        self.max_cut = float((self.table[..., 10] + self.table[..., 11]).max())


def _ters_fc(r, R, D):
    x = (r - R) / D
    mid = 0.5 * (1.0 - torch.sin(0.5 * math.pi * x.clamp(-1.0, 1.0)))
    return torch.where(r < R - D, torch.ones_like(r), torch.where(r > R + D, torch.zeros_like(r), mid))


def tersoff_energy(pos: torch.Tensor, types: torch.Tensor, cell, pbc, params: TersoffParams,
                   nbrs=None) -> torch.Tensor:
    """Total Tersoff energy (eV), differentiable w.r.t. ``pos`` (fp64)."""
    if nbrs is None:
        i, j, S = neighbor_list(pos.detach().numpy(), cell, pbc, params.max_cut + 1e-3)
    else:
        i, j, S = nbrs
    off = torch.tensor(S.astype(np.float64) @ np.asarray(cell, dtype=np.float64))
    i = torch.as_tensor(i, dtype=torch.long)
    j = torch.as_tensor(j, dtype=torch.long)
    tab = torch.tensor(params.table)
    rij_vec = pos[j] - pos[i] + off
    rij = rij_vec.norm(dim=1)
    ti, tj = types[i], types[j]
    pij = tab[ti, tj, tj]                       # [E,14]
    cut_ij = pij[:, 10] + pij[:, 11]
    keep = rij < cut_ij                          # LAMMPS: rsq >= cutsq -> skip
    i, j, rij_vec, rij, ti, tj, pij = i[keep], j[keep], rij_vec[keep], rij[keep], ti[keep], tj[keep], pij[keep]
    n_e = i.shape[0]
    # triplets: for each edge e=(i,j) all edges e2=(i,k) with e2 != e
    order = torch.argsort(i, stable=True)
    assert torch.equal(order, torch.arange(n_e)), "edges must be receiver-sorted"
    n_atoms = pos.shape[0]
    cnt = torch.bincount(i, minlength=n_atoms)
    start = torch.cumsum(cnt, 0) - cnt
    e1 = torch.repeat_interleave(torch.arange(n_e), cnt[i])
    # local index of the partner edge inside the row of atom i
    within = torch.arange(e1.shape[0]) - torch.repeat_interleave(torch.cumsum(cnt[i], 0) - cnt[i], cnt[i])
    e2 = start[i[e1]] + within
    m = e1 != e2
    e1, e2 = e1[m], e2[m]
    tk = tj[e2]
    pijk = tab[ti[e1], tj[e1], tk]               # [T,14]
    rik = rij[e2]
    keep3 = rik < (pijk[:, 10] + pijk[:, 11])
    e1, e2, pijk, rik = e1[keep3], e2[keep3], pijk[keep3], rik[keep3]
    cos_t = (rij_vec[e1] * rij_vec[e2]).sum(1) / (rij[e1] * rik)
    fc_ik = _ters_fc(rik, pijk[:, 10], pijk[:, 11])
    c2, d2 = pijk[:, 3] ** 2, pijk[:, 4] ** 2
    hcth = pijk[:, 5] - cos_t
    g = pijk[:, 1] * (1.0 + c2 / d2 - c2 / (d2 + hcth * hcth))
    arg = pijk[:, 2] * (rij[e1] - rik)
    arg = torch.where(pijk[:, 0] == 3.0, arg ** 3, arg)
    ex = torch.where(arg > 69.0776, torch.full_like(arg, 1e30),
                     torch.where(arg < -69.0776, torch.zeros_like(arg), torch.exp(arg.clamp(-69.0776, 69.0776))))
    zeta = torch.zeros(n_e, dtype=pos.dtype).index_add_(0, e1, fc_ik * g * ex)
    # b_ij with the LAMMPS asymptotic branches
    n_p, beta = pij[:, 6], pij[:, 7]
    tmp = beta * zeta
    c1 = (2.0 * n_p * 1.0e-16) ** (-1.0 / n_p)
    c2b = (2.0 * n_p * 1.0e-8) ** (-1.0 / n_p)
    c3 = 1.0 / c2b
    c4 = 1.0 / c1
    safe = tmp.clamp_min(1e-300)
    b_mid = (1.0 + safe ** n_p) ** (-1.0 / (2.0 * n_p))
    bij = torch.where(tmp > c1, 1.0 / safe.sqrt(),
          torch.where(tmp > c2b, (1.0 - safe ** (-n_p) / (2.0 * n_p)) / safe.sqrt(),
          torch.where(tmp < c4, torch.ones_like(tmp),
          torch.where(tmp < c3, 1.0 - safe ** n_p / (2.0 * n_p), b_mid))))
    fc_ij = _ters_fc(rij, pij[:, 10], pij[:, 11])
    e_rep = fc_ij * pij[:, 13] * torch.exp(-pij[:, 12] * rij)
    e_att = -bij * pij[:, 9] * torch.exp(-pij[:, 8] * rij) * fc_ij
    return 0.5 * (e_rep + e_att).sum()


@dataclass
class SWParams:
    """LAMMPS ``pair_style sw`` single-element parameter struct (SW-1985 Si defaults)."""
    epsilon: float = 2.1683
    sigma: float = 2.0951
    a: float = 1.80
    lam: float = 21.0
    gamma: float = 1.20
    costheta0: float = -1.0 / 3.0
    A: float = 7.049556277
    B: float = 0.6022245584
    p: float = 4.0
    q: float = 0.0

    @property
    def cut(self) -> float:
        return self.a * self.sigma


def sw_energy(pos: torch.Tensor, cell, pbc, prm: SWParams, nbrs=None) -> torch.Tensor:
    """Total Stillinger-Weber energy (eV), differentiable w.r.t. ``pos`` (fp64)."""
    if nbrs is None:
        i, j, S = neighbor_list(pos.detach().numpy(), cell, pbc, prm.cut + 1e-3)
    else:
        i, j, S = nbrs
    off = torch.tensor(S.astype(np.float64) @ np.asarray(cell, dtype=np.float64))
    i = torch.as_tensor(i, dtype=torch.long)
    j = torch.as_tensor(j, dtype=torch.long)
    rv = pos[j] - pos[i] + off
    r = rv.norm(dim=1)
    keep = r < prm.cut
    i, j, rv, r = i[keep], j[keep], rv[keep], r[keep]
    n_e = i.shape[0]
    sr = prm.sigma / r
    e2 = prm.A * prm.epsilon * (prm.B * sr ** prm.p - sr ** prm.q) * torch.exp(prm.sigma / (r - prm.cut))
    n_atoms = pos.shape[0]
    cnt = torch.bincount(i, minlength=n_atoms)
    start = torch.cumsum(cnt, 0) - cnt
    ea = torch.repeat_interleave(torch.arange(n_e), cnt[i])
    within = torch.arange(ea.shape[0]) - torch.repeat_interleave(torch.cumsum(cnt[i], 0) - cnt[i], cnt[i])
    eb = start[i[ea]] + within
    m = ea < eb   # j<k: each unordered pair of bonds once
    ea, eb = ea[m], eb[m]
    cos_t = (rv[ea] * rv[eb]).sum(1) / (r[ea] * r[eb])
    e3 = prm.lam * prm.epsilon * (cos_t - prm.costheta0) ** 2 \
        * torch.exp(prm.gamma * prm.sigma / (r[ea] - prm.cut)) \
        * torch.exp(prm.gamma * prm.sigma / (r[eb] - prm.cut))
    return 0.5 * e2.sum() + e3.sum()


def energy_forces(fn, pos_np, *args, **kw):
    """(E, F) from a differentiable energy function."""
    pos = torch.tensor(np.asarray(pos_np, dtype=np.float64), requires_grad=True)
    e = fn(pos, *args, **kw)
    (g,) = torch.autograd.grad(e, pos)
    return float(e.detach()), -g.numpy()
