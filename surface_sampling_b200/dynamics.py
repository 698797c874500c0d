"""``optimize_slab`` — drop-in for the reference's relax call (mcmc/dynamics.py:83-170).

Same signature and return tuple ``(calc_slab, traj, energy, energy_oob)``; same thresholds
(ENERGY_THRESHOLD / MAX_FORCE_THRESHOLD = 1000, dynamics.py:16-17).  FIRE runs on the GPU:
  * ``save_traj=False``  -> one fused call (vssr_painn_relax / vssr_classical_relax), nothing but the
    final positions and 8 scalars per structure come back;
  * ``save_traj=True``   -> the same kernels driven step by step so the observer can record every
    ``record_interval`` steps (``traj = {"atoms", "energies", "forces"}`` as in dynamics.py:145-150).
Only ``optimizer="FIRE"`` (the default) and ``"LAMMPS"`` are served by the engine; BFGS/CG are
host-side numpy optimisers in the reference and are out of the hot path.
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
    if not hasattr(engine, "set_framework") or not np.any(fixed):
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
        if optimizer not in ("FIRE",):
            raise NotImplementedError(f"optimizer {optimizer!r}: only FIRE and LAMMPS run on the B200 engine")
        relax_steps = kwargs.get("relax_steps", 20)
        record_interval = kwargs.get("record_interval", 5)
        engine = calc.engine
        pos, num, cell, pbc, fixed = as_arrays(slab)
        batch = eng.Batch.from_arrays([pos], [num], [cell], [pbc], [fixed])
        _auto_framework(engine, pos, cell, pbc, fixed)
        calc_slab = slab.copy()
        calc_slab.calc = calc
        if not save_traj:
            r = engine.relax(batch, relax_steps=relax_steps, fmax=0.01, z_host=num, check=True)
            out = r["out"].cpu().numpy()[0]
            energy = float(out[2])
            forces = r["forces"].cpu().numpy()
            traj = None
        else:
            obs = {"atoms": [], "energies": [], "forces": []}

            def observer(step, e, f, p):
                if step % record_interval == 0:
                    a = slab.copy()
                    a.calc = None
                    _set_positions(a, p)
                    obs["atoms"].append(a)
                    obs["energies"].append(float(e))
                    fm = np.array(f, dtype=float)
                    fm[fixed] = 0.0
                    obs["forces"].append(fm)

            energy, forces = relax_stepwise(engine, batch, num, relax_steps, 0.01, observer)
            traj = obs
        _set_positions(calc_slab, batch.pos.cpu().numpy())
        # leave the calculator primed with the final evaluation like ASE does
        calc.results.update({"energy": np.array([energy], dtype=np.float32), "forces": forces})
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
