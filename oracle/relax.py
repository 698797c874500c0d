"""ORACLE (test infrastructure only — never imported by the product path).

CPU restatement (numpy fp64) of the relaxation driver on the reference's hot path:
``optimize_slab`` (mcmc/dynamics.py:83-170) = ASE optimizer ``run(steps=relax_steps, fmax=0.01)``
+ the out-of-bounds clamp (ENERGY_THRESHOLD / MAX_FORCE_THRESHOLD = 1000, dynamics.py:16-17,
159-168).  ASE (``ase>=3.22.1,<=3.23.0``, pyproject.toml:14) is un-vendored; FIRE, BFGS, the
``Dynamics.irun`` loop and ``FixAtoms`` are restated from its published algorithm
(SURVEY.md App. A.5).

Pinning: the BFGS path reproduces the 5-line log of the pristine SrTiO3 slab
(tutorials/SrTiO3_001.ipynb:241-245) step for step -> pins the driver loop, FixAtoms masking,
fmax test and the ensemble calculator together (tests/test_oracle_golden.py).  FIRE itself has
no log anywhere in the reference tree: FIRE TRAJECTORY PARITY IS UNPINNED in-tree; the
surrounding driver semantics are pinned through BFGS.

Convention chosen for FIRE (documented in DESIGN.md): forces arrive in the calculator's dtype
(fp32 for PaiNN), are cast to fp64, and all optimizer arithmetic is fp64.
"""
from __future__ import annotations

import numpy as np

ENERGY_THRESHOLD = 1000.0
MAX_FORCE_THRESHOLD = 1000.0


class FIRE:
    """ASE FIRE defaults: dt=0.1, maxstep=0.2, dtmax=1.0, Nmin=5, finc=1.1, fdec=0.5,
    astart=0.1, fa=0.99; no masses; global (whole-structure) norms."""

    def __init__(self, n_atoms, dt=0.1, maxstep=0.2, dtmax=1.0, nmin=5, finc=1.1, fdec=0.5,
                 astart=0.1, fa=0.99):
        self.dt, self.maxstep, self.dtmax = dt, maxstep, dtmax
        self.nmin, self.finc, self.fdec, self.astart, self.fa = nmin, finc, fdec, astart, fa
        self.a = astart
        self.v = None
        self.nsteps_pos = 0
        self.n = n_atoms

    def step(self, x, f):
        f = np.asarray(f, dtype=np.float64)
        if self.v is None:
            self.v = np.zeros((self.n, 3))
        else:
            vf = np.vdot(f, self.v)
            if vf > 0.0:
                self.v = (1.0 - self.a) * self.v + self.a * f / np.sqrt(np.vdot(f, f)) * np.sqrt(
                    np.vdot(self.v, self.v))
                if self.nsteps_pos > self.nmin:
                    self.dt = min(self.dt * self.finc, self.dtmax)
                    self.a *= self.fa
                self.nsteps_pos += 1
            else:
                self.v[:] *= 0.0
                self.a = self.astart
                self.dt *= self.fdec
                self.nsteps_pos = 0
        self.v += self.dt * f
        dr = self.dt * self.v
        normdr = np.sqrt(np.vdot(dr, dr))
        if normdr > self.maxstep:
            dr = self.maxstep * dr / normdr
        return x + dr


class BFGS:
    """ASE BFGS: H0 = 70 I, maxstep 0.2, eigen-decomposition step."""

    def __init__(self, n_atoms, alpha=70.0, maxstep=0.2):
        self.H = None
        self.alpha = alpha
        self.maxstep = maxstep
        self.r0 = None
        self.f0 = None
        self.n = n_atoms

    def step(self, x, f):
        f = np.asarray(f, dtype=np.float64).reshape(-1)
        r = np.asarray(x, dtype=np.float64).reshape(-1)
        if self.H is None:
            self.H = np.eye(3 * self.n) * self.alpha
        else:
            dr = r - self.r0
            if np.abs(dr).max() >= 1e-7:
                df = f - self.f0
                a = np.dot(dr, df)
                dg = np.dot(self.H, dr)
                b = np.dot(dr, dg)
                self.H -= np.outer(df, df) / a + np.outer(dg, dg) / b
        omega, V = np.linalg.eigh(self.H)
        dx = np.dot(V, np.dot(f, V) / np.fabs(omega)).reshape(-1, 3)
        steplengths = (dx ** 2).sum(1) ** 0.5
        maxsl = steplengths.max()
        if maxsl >= self.maxstep:
            dx *= self.maxstep / maxsl
        self.r0 = r.copy()
        self.f0 = f.copy()
        return np.asarray(x, dtype=np.float64) + dx


def relax(calc_fn, pos, fixed_mask, optimizer="FIRE", relax_steps=20, fmax=0.01, log=None):
    """``Optimizer.run(steps, fmax)`` restated: evaluate; while max_i |F_i| >= fmax and
    nsteps < steps: step(); evaluate.  <= steps+1 evaluations.

    Args:
        calc_fn: pos[N,3] fp64 -> (energy float, forces [N,3]) (raw, un-constrained).
        fixed_mask: [N] bool, True = FixAtoms (force zeroed, position frozen).
    Returns: dict(pos, energy, forces_raw, nsteps, converged, energy_oob, n_evals)
    """
    x = np.array(pos, dtype=np.float64)
    n = x.shape[0]
    fixed_mask = np.asarray(fixed_mask, dtype=bool)
    opt = FIRE(n) if optimizer == "FIRE" else BFGS(n)
    nsteps = 0
    n_evals = 0
    while True:
        e, f_raw = calc_fn(x)
        n_evals += 1
        f = np.array(f_raw, dtype=np.float64)
        f[fixed_mask] = 0.0
        fm = np.sqrt((f ** 2).sum(1).max()) if n else 0.0
        if log is not None:
            log.append((nsteps, float(e), float(fm)))
        converged = (f ** 2).sum(1).max() < fmax ** 2
        if converged or nsteps >= relax_steps:
            break
        x_new = opt.step(x, f)
        x_new[fixed_mask] = x[fixed_mask]
        x = x_new
        nsteps += 1
    energy = float(e)
    max_force = float(np.abs(np.asarray(f_raw)).max()) if n else 0.0
    oob = bool(abs(energy) > ENERGY_THRESHOLD or max_force > MAX_FORCE_THRESHOLD)
    return {
        "pos": x, "energy": ENERGY_THRESHOLD if oob else energy, "raw_energy": energy,
        "forces_raw": np.asarray(f_raw), "nsteps": nsteps, "converged": bool(converged),
        "energy_oob": oob, "n_evals": n_evals,
    }


def catkit_layer_tags(pos, cell) -> np.ndarray:
    """CatKit ``get_unique_coordinates(tag=True)`` restated (free-atom set):
    greedy clusters of scaled z under isclose(atol=1e-3, rtol=1e-3); tag 1 = topmost layer."""
    cell = np.asarray(cell, dtype=np.float64)
    z = (np.asarray(pos, dtype=np.float64) @ np.linalg.inv(cell))[:, 2]
    z = np.round(z % 1.0, 4) if False else z
    reps: list[float] = []
    layer = np.zeros(len(z), dtype=int)
    for a, zz in enumerate(z):
        for k, r in enumerate(reps):
            if np.isclose(zz, r, atol=1e-3, rtol=1e-3):
                layer[a] = k
                break
        else:
            reps.append(zz)
            layer[a] = len(reps) - 1
    order = np.argsort(-np.array(reps))  # topmost first
    rank = np.empty(len(reps), dtype=int)
    rank[order] = np.arange(1, len(reps) + 1)
    return rank[layer]


def fixed_mask_from_surface_depth(pos, cell, surface_depth: int) -> np.ndarray:
    """SurfaceSystem.initialize_constraints (mcmc/system.py:268-300): free tags 1..depth."""
    tags = catkit_layer_tags(pos, cell)
    return ~np.isin(tags, list(range(1, surface_depth + 1)))
