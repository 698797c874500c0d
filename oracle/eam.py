"""ORACLE (test infrastructure only — never imported by the product path).

CPU restatement (numpy fp64) of LAMMPS ``pair_style eam`` with a single-element ``funcfl`` file,
the potential behind the reference's Cu/Au toy runs (BASELINE config 1):
``LAMMPSRunSurfCalc`` (mcmc/calculators/calculators.py:755-811) -> forked ASE LAMMPS runner
(mcmc/calculators/lammpsrun.py:309-469) -> ``lmp`` subprocess with ``mcmc/potentials/{Cu,Au}_u3.eam``.
LAMMPS is an un-vendored, unpinned conda dependency (environment.yml:5-8); the algorithm follows
its published pair_eam.cpp:

  * file: line 2 = Z, mass, a0, lattice; line 3 = Nrho drho Nr dr rc; then F(rho)[Nrho], Z(r)[Nr],
    rho(r)[Nr];
  * phi(r) = 27.2 * 0.529 * Z(r)^2 / r   (tabulated as z2r = r*phi, divided by r at lookup);
  * E = sum_i F(sum_j rho(r_ij)) + 1/2 sum_ij phi(r_ij);
  * every table becomes LAMMPS' 7-coefficient cubic spline (``interpolate``: end slopes by
    differences, interior 5-point stencil ((f[m-2]-f[m+2]) + 8(f[m+1]-f[m-1]))/12), looked up with
    p = x/dx + 1; m = int(p) clipped to [1, n-1]; p -= m; p = min(p, 1).

PINNED: tests/test_Au.py:19 of the reference asserts min(energy_hist) = -79.03490823689619 for the
canonical Au(110) run whose state space is the 28 ways of keeping 6 of the 8 adatoms of
tests/data/Au_110/Au_110_2x2_proper_adsorbed_slab.cif (tests/test_oracle_golden.py).
"""
from __future__ import annotations

import numpy as np

from .nbrlist import neighbor_list


def _spline(f: np.ndarray, delta: float) -> np.ndarray:
    """LAMMPS PairEAM::interpolate -> spline[m][0..6] for m = 1..n (index 0 unused)."""
    n = len(f)
    s = np.zeros((n + 1, 7))
    s[1:, 6] = f
    s[1, 5] = s[2, 6] - s[1, 6]
    s[2, 5] = 0.5 * (s[3, 6] - s[1, 6])
    s[n - 1, 5] = 0.5 * (s[n, 6] - s[n - 2, 6])
    s[n, 5] = s[n, 6] - s[n - 1, 6]
    m = np.arange(3, n - 1)
    s[m, 5] = ((s[m - 2, 6] - s[m + 2, 6]) + 8.0 * (s[m + 1, 6] - s[m - 1, 6])) / 12.0
    m = np.arange(1, n)
    s[m, 4] = 3.0 * (s[m + 1, 6] - s[m, 6]) - 2.0 * s[m, 5] - s[m + 1, 5]
    s[m, 3] = s[m, 5] + s[m + 1, 5] - 2.0 * (s[m + 1, 6] - s[m, 6])
    s[n, 4] = 0.0
    s[n, 3] = 0.0
    s[1:, 2] = s[1:, 5] / delta
    s[1:, 1] = 2.0 * s[1:, 4] / delta
    s[1:, 0] = 3.0 * s[1:, 3] / delta
    return s


def _lookup(spl: np.ndarray, x: np.ndarray, rdx: float, n: int):
    p = x * rdx + 1.0
    m = np.clip(p.astype(np.int64), 1, n - 1)
    p = np.minimum(p - m, 1.0)
    c = spl[m]
    val = ((c[:, 3] * p + c[:, 4]) * p + c[:, 5]) * p + c[:, 6]
    der = (c[:, 0] * p + c[:, 1]) * p + c[:, 2]
    return val, der


class EAMFuncfl:
    def __init__(self, tab: dict):
        self.nrho, self.drho = int(tab["nrho"]), float(tab["drho"])
        self.nr, self.dr, self.rc = int(tab["nr"]), float(tab["dr"]), float(tab["rc"])
        frho = np.asarray(tab["frho"], dtype=np.float64)
        zr = np.asarray(tab["zr"], dtype=np.float64)
        rhor = np.asarray(tab["rhor"], dtype=np.float64)
        r = np.arange(self.nr) * self.dr
        z2r = 27.2 * 0.529 * zr * zr          # = r * phi(r)
        self.frho_spl = _spline(frho, self.drho)
        self.rhor_spl = _spline(rhor, self.dr)
        self.z2r_spl = _spline(z2r, self.dr)
        self.rhomax = (self.nrho - 1) * self.drho

    def energy_forces(self, pos, cell, pbc):
        pos = np.asarray(pos, dtype=np.float64)
        i, j, S = neighbor_list(pos, cell, pbc, self.rc)
        rv = pos[j] - pos[i] + S.astype(np.float64) @ np.asarray(cell, dtype=np.float64)
        r = np.linalg.norm(rv, axis=1)
        keep = r < self.rc
        i, j, rv, r = i[keep], j[keep], rv[keep], r[keep]
        n = len(pos)
        rho_e, drho_e = _lookup(self.rhor_spl, r, 1.0 / self.dr, self.nr)
        rho = np.bincount(i, weights=rho_e, minlength=n)
        F, dF = _lookup(self.frho_spl, rho, 1.0 / self.drho, self.nrho)
        over = rho > self.rhomax                      # LAMMPS linear extrapolation beyond the table
        F = np.where(over, F + dF * (rho - self.rhomax), F)
        z2, dz2 = _lookup(self.z2r_spl, r, 1.0 / self.dr, self.nr)
        phi = z2 / r
        dphi = dz2 / r - phi / r
        energy = F.sum() + 0.5 * phi.sum()
        # dE/dr_ij for the directed edge (i<-j): 1/2 phi' counted twice over both directions
        fpair = (dF[i] + dF[j]) * drho_e * 0.5 + 0.5 * dphi
        g = fpair[:, None] * rv / r[:, None]          # dE/d(x_j) contribution; -g on x_i
        forces = np.zeros((n, 3))
        np.add.at(forces, i, g)
        np.add.at(forces, j, -g)
        return float(energy), forces
