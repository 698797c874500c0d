"""``optimize_slab`` — drop-in for the reference's relax call (mcmc/dynamics.py:83-170).

Same signature and return tuple ``(calc_slab, traj, energy, energy_oob)``; same thresholds
(ENERGY_THRESHOLD / MAX_FORCE_THRESHOLD = 1000, dynamics.py:16-17).  FIRE runs on the GPU:
  * ``save_traj=False``  -> one fused call (vssr_painn_relax / vssr_classical_relax), nothing but the
    final positions and 8 scalars per structure come back;
  * ``save_traj=True``   -> the same kernels driven step by step so the observer can record every
    ``record_interval`` steps (``traj = {"atoms", "energies", "forces"}`` as in dynamics.py:145-150).
``optimizer="BFGS"`` (what every PaiNN notebook of the reference uses, tutorials/SrTiO3_001.ipynb:241-245) and
``"CG"`` are host-side optimisers in the reference as well (ASE numpy / scipy code around the calculator); here the
host keeps only the quasi-Newton bookkeeping (one small ``eigh`` per structure and step, restricted to the free
atoms) while every energy/force evaluation is the batched CUDA ensemble call (``relax_host_batch``).
``"BFGSLineSearch"`` is not provided (ASE's More-Thuente line search is not restated here).
"""
from __future__ import annotations

import logging

import numpy as np
import torch

from . import _lib
from . import engine as eng
from .atoms import as_arrays

ENERGY_THRESHOLD = 1000  # eV
MAX_FORCE_THRESHOLD = 1000  # eV/Angstrom


def _set_positions(atoms, pos):
    if hasattr(atoms, "set_positions"):
        try:
            atoms.set_positions(pos, apply_constraint=False)
            return
        except TypeError:
            pass
    atoms.positions = np.asarray(pos, dtype=float).copy()


def _auto_framework(engine, pos, cell, pbc, fixed):
    """Register the slab's FixAtoms block as the engine's frozen framework (radial-filter memo + no dE/dx on
    frozen atoms during the FIRE steps; engine.PainnEngine.set_framework).  The MC loop calls optimize_slab
    with the same bulk over and over, so the one-time build is cached on the engine by a hash of the frozen
    coordinates.  Results of the relaxation are unchanged (tests/test_gpu_painn.py, test_gpu_boundary.py)."""
    if not hasattr(engine, "set_framework"):
        return
    if not np.any(fixed):
        # a slab without FixAtoms must not inherit the previous slab's frozen framework (the engine would reject
        # the batch: "a structure does not hold the framework's frozen atoms fixed")
        if getattr(engine, "_auto_fc_key", None) is not None:
            engine.clear_framework()
            engine._auto_fc_key = None
        return
    n0 = int(np.flatnonzero(fixed)[-1]) + 1
    p0 = np.ascontiguousarray(pos[:n0], dtype=np.float32)
    key = (n0, hash(p0[np.asarray(fixed[:n0], bool)].tobytes()), hash(np.asarray(cell, dtype=np.float64).tobytes()),
           hash(np.asarray(fixed[:n0], bool).tobytes()))
    if getattr(engine, "_auto_fc_key", None) != key:
        engine.set_framework(p0, cell, pbc, np.asarray(fixed[:n0], bool), constrained_forces=True)
        engine._auto_fc_key = key


def optimize_slab(slab, optimizer="FIRE", save_traj=True, logger=None, **kwargs) -> tuple:
    logger = logger or logging.getLogger(__name__)
    calc = slab.calc
    if "LAMMPS" in optimizer:
        calc_slab, energy, _ = calc.run_lammps_opt(slab, run_dir=calc.run_dir,
                                                   **{k: v for k, v in kwargs.items() if k == "relax_steps"})
        traj = None
        max_force = float(np.abs(getattr(calc, "_last_forces", np.zeros(1))).max())
    else:
        if "BFGSLineSearch" in optimizer:
            raise NotImplementedError("optimizer 'BFGSLineSearch' is not provided: use FIRE, BFGS, CG or LAMMPS")
        host_opt = "BFGS" if "BFGS" in optimizer else ("CG" if "CG" in optimizer else None)   # dynamics.py:120-127
        relax_steps = kwargs.get("relax_steps", 20)
        record_interval = kwargs.get("record_interval", 5)
        engine = calc.engine
        pos, num, cell, pbc, fixed = as_arrays(slab)
        batch = eng.Batch.from_arrays([pos], [num], [cell], [pbc], [fixed])
        _auto_framework(engine, pos, cell, pbc, fixed)
        calc_slab = slab.copy()
        calc_slab.calc = calc
        obs = {"atoms": [], "energies": [], "forces": []} if save_traj else None

        def observer(step, e, f, p):     # TrajectoryObserver (dynamics.py:20-80), attached at `record_interval`
            if obs is not None and step % record_interval == 0:
                a = slab.copy()
                a.calc = None
                _set_positions(a, p)
                obs["atoms"].append(a)
                obs["energies"].append(float(e))
                fm = np.array(f, dtype=float)
                fm[fixed] = 0.0
                obs["forces"].append(fm)

        if host_opt is not None:
            out, forces = relax_host_batch(engine, batch, num, host_opt, relax_steps, 0.01, observer=observer)
            energy = float(out[0, 2])
        elif not save_traj:
            r = engine.relax(batch, relax_steps=relax_steps, fmax=0.01, z_host=num, check=True)
            energy = float(r["out"].cpu().numpy()[0, 2])
            forces = r["forces"].cpu().numpy()
        else:
            energy, forces = relax_stepwise(engine, batch, num, relax_steps, 0.01, observer)
        traj = obs
        _set_positions(calc_slab, batch.pos.cpu().numpy())
        # leave the calculator primed with the final evaluation like ASE does
        # (REPLACE the results: surface_energy / energy_std / embedding of an earlier structure must not survive)
        calc.results = {"energy": np.array([energy], dtype=np.float32), "forces": forces}
        calc._cache_key = calc._key(calc_slab)
        max_force = float(np.abs(forces).max())

    if np.abs(energy) > ENERGY_THRESHOLD or max_force > MAX_FORCE_THRESHOLD:
        logger.info("encountered energy or force out of bounds")
        logger.info("energy %.3f", energy)
        logger.info("max force %.3f", max_force)
        energy = ENERGY_THRESHOLD
        energy_oob = True
    else:
        energy_oob = False
    return calc_slab, traj, energy, energy_oob


def relax_stepwise(engine: "eng.PainnEngine", batch: "eng.Batch", z_host, relax_steps, fmax, observer=None):
    """Host-driven FIRE for a batch of ONE structure using the same kernels as the fused path
    (energy_forces + vssr_fire_step); exists for the trajectory observer."""
    lib = _lib.load()
    dev = batch.pos.device
    A, B = batch.n_atoms, batch.n_struct
    assert B == 1
    nbrs = eng.neighbor_list(batch, engine.cutoff + engine.skin)
    state = torch.empty((B, 8), dtype=torch.float64, device=dev)
    vel = torch.empty((A, 3), dtype=torch.float64, device=dev)
    pos32 = batch.pos.to(torch.float32)
    _lib.check(lib.vssr_fire_init(state.data_ptr(), vel.data_ptr(), B, A, eng._stream()), "vssr_fire_init")
    energy, forces = None, None
    for it in range(relax_steps + 1):
        r = engine.energy_forces(batch, z_host=z_host, nbrs=nbrs)
        energy = float(r["energy"].item())
        forces = r["forces"].cpu().numpy()
        if observer is not None:
            observer(it, energy, forces, batch.pos.cpu().numpy())
        _lib.check(lib.vssr_fire_step(batch.pos.data_ptr(), pos32.data_ptr(), vel.data_ptr(), r["forces"].data_ptr(),
                                      batch.fixed.data_ptr(), batch.atom_ptr.data_ptr(), B, state.data_ptr(),
                                      int(relax_steps), float(fmax), eng._stream()), "vssr_fire_step")
        if state[0, 4].item() != 0.0:   # converged: ASE stops calling the calculator
            break
    return energy, forces


class HostBFGS:
    """ASE ``BFGS`` (un-vendored dependency) for one structure, kept in the subspace of the free
    atoms: with FixAtoms the force rows of fixed atoms are zero, so the 70*I block of their coordinates is never
    updated and never contributes to a step -- the restriction changes nothing but the size of the eigenproblem."""

    def __init__(self, free_idx, alpha=70.0, maxstep=0.2):
        self.free = np.asarray(free_idx, dtype=int)
        self.alpha, self.maxstep = float(alpha), float(maxstep)
        self.H = self.r_prev = self.f_prev = None

    def step(self, x, f):
        """x, f: [N,3] fp64 positions / constraint-masked forces. Returns the displacement [N,3]."""
        r = x[self.free].reshape(-1)
        g = f[self.free].reshape(-1)
        if self.H is None:
            self.H = np.eye(r.size) * self.alpha
        else:
            dr = r - self.r_prev
            if np.abs(dr).max() >= 1e-7:
                df = g - self.f_prev
                Hdr = self.H @ dr
                self.H -= np.outer(df, df) / (dr @ df) + np.outer(Hdr, Hdr) / (dr @ Hdr)
        omega, V = np.linalg.eigh(self.H)
        step = (V @ ((g @ V) / np.fabs(omega))).reshape(-1, 3)
        longest = np.sqrt((step ** 2).sum(1).max()) if len(step) else 0.0
        if longest >= self.maxstep:
            step *= self.maxstep / longest
        self.r_prev, self.f_prev = r.copy(), g.copy()
        dx = np.zeros_like(x)
        dx[self.free] = step
        return dx


def relax_host_batch(engine: "eng.PainnEngine", batch: "eng.Batch", z_host, optimizer="BFGS", relax_steps=20, fmax=0.01,
                     observer=None):
    """``Optimizer(atoms).run(steps=relax_steps, fmax=fmax)`` (mcmc/dynamics.py:133-141) for every structure of the batch
    with a host-side optimiser: ``Dynamics.irun`` = evaluate; while max_i|F_i| >= fmax and nsteps < steps: step, evaluate.
    One batched ensemble evaluation on the GPU per optimiser step serves all structures; a structure that has converged
    simply stops moving.  batch.pos is updated in place.  Returns (out[B,8] like vssr_painn_relax, forces[A,3] raw fp32
    of each structure's LAST evaluation)."""
    if optimizer == "CG":
        if batch.n_struct != 1:
            raise NotImplementedError("optimizer 'CG' (scipy fmin_cg) is served one structure at a time")
        return _relax_host_cg(engine, batch, z_host, relax_steps, fmax, observer)
    B, ptr = batch.n_struct, batch.atom_ptr_host
    fixed = batch.fixed_host if batch.fixed_host is not None else batch.fixed.cpu().numpy().astype(bool)
    nbrs = eng.neighbor_list(batch, engine.cutoff + engine.skin)     # once per relaxation (dynamics.py:129)
    pos = batch.pos.cpu().numpy().copy()
    opts = [HostBFGS(np.flatnonzero(~fixed[ptr[b]:ptr[b + 1]])) for b in range(B)]
    out = np.zeros((B, 8))
    forces_last = np.zeros((batch.n_atoms, 3), dtype=np.float32)
    done = np.zeros(B, dtype=bool)
    nsteps = np.zeros(B, dtype=int)
    for _ in range(relax_steps + 1):
        r = engine.energy_forces(batch, z_host=z_host, nbrs=nbrs)
        e, es, f = r["energy"].cpu().numpy(), r["energy_std"].cpu().numpy(), r["forces"].cpu().numpy()
        moved = False
        for b in range(B):
            if done[b]:
                continue
            lo, hi = ptr[b], ptr[b + 1]
            fb = f[lo:hi].astype(np.float64)
            fm = fb.copy()
            fm[fixed[lo:hi]] = 0.0
            f2 = (fm ** 2).sum(1).max() if hi > lo else 0.0
            conv = f2 < fmax ** 2
            out[b, :4] = (e[b], es[b], e[b], np.abs(fb).max() if hi > lo else 0.0)
            out[b, 4], out[b, 5], out[b, 7] = nsteps[b], float(conv), nsteps[b] + 1
            forces_last[lo:hi] = f[lo:hi]
            if observer is not None and B == 1:
                observer(int(nsteps[b]), e[b], f[lo:hi], pos[lo:hi])
            if conv or nsteps[b] >= relax_steps:
                done[b] = True
                continue
            pos[lo:hi] += opts[b].step(pos[lo:hi], fm)
            nsteps[b] += 1
            moved = True
        if not moved:
            break
        batch.pos.copy_(torch.from_numpy(pos))
    oob = (np.abs(out[:, 2]) > ENERGY_THRESHOLD) | (out[:, 3] > MAX_FORCE_THRESHOLD)
    out[oob, 0], out[:, 6] = ENERGY_THRESHOLD, oob
    return out, forces_last


def _relax_host_cg(engine, batch, z_host, relax_steps, fmax, observer):
    """ASE ``SciPyFminCG`` (ase.optimize.sciopt): scipy's Polak-Ribiere ``fmin_cg`` on E/alpha with alpha = 70,
    gtol = 0.1*fmax/alpha in the max-norm, maxiter = steps, convergence checked in the per-iteration callback."""
    from scipy import optimize as sopt
    alpha = 70.0
    fixed = batch.fixed_host if batch.fixed_host is not None else batch.fixed.cpu().numpy().astype(bool)
    nbrs = eng.neighbor_list(batch, engine.cutoff + engine.skin)
    x_fixed = batch.pos.cpu().numpy().copy()
    cache = {}

    class Converged(Exception):
        pass

    def evaluate(x):
        p = x.reshape(-1, 3).copy()
        p[fixed] = x_fixed[fixed]                       # Atoms.set_positions applies FixAtoms
        key = p.tobytes()
        if cache.get("key") != key:
            batch.pos.copy_(torch.from_numpy(p))
            r = engine.energy_forces(batch, z_host=z_host, nbrs=nbrs)
            cache.update(key=key, e=float(r["energy"].item()), es=float(r["energy_std"].item()),
                         f=r["forces"].cpu().numpy(), p=p, n=cache.get("n", 0) + 1)
        return cache

    def masked(c):
        fm = c["f"].astype(np.float64)
        fm[fixed] = 0.0
        return fm

    state = {"nsteps": 0}

    def callback(x):
        c = evaluate(cache["p"].reshape(-1) if x is None else x)
        if observer is not None:
            observer(state["nsteps"], c["e"], c["f"], c["p"])
        if (masked(c) ** 2).sum(1).max() < fmax ** 2:
            raise Converged
        state["nsteps"] += 1

    evaluate(x_fixed.reshape(-1))
    conv = False
    try:
        callback(None)
        sopt.fmin_cg(lambda x: evaluate(x)["e"] / alpha, x_fixed.reshape(-1), fprime=lambda x: -masked(evaluate(x)).reshape(-1) / alpha,
                     gtol=fmax / alpha * 0.1, norm=np.inf, maxiter=relax_steps, full_output=1, disp=0, callback=callback)
    except Converged:
        conv = True
    c = cache
    batch.pos.copy_(torch.from_numpy(c["p"]))
    maxf = float(np.abs(c["f"]).max())
    oob = abs(c["e"]) > ENERGY_THRESHOLD or maxf > MAX_FORCE_THRESHOLD
    out = np.array([[ENERGY_THRESHOLD if oob else c["e"], c["es"], c["e"], maxf, state["nsteps"], float(conv), float(oob), c["n"]]])
    return out, c["f"]
