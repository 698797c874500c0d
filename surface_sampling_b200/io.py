"""Per-sweep structure dumps of the MC run: what ``SurfaceSystem.save_structures``
(mcmc/system.py:488-534) writes after every sweep, without ASE.

  {oob}_unrelaxed_slab_sweep_{NNN}_energy_{E:.3f}_{formula}.cif     the MC state (ideal-site structure)
  {oob}_relaxed_slab_sweep_{NNN}_energy_{E:.3f}_{formula}.cif       the relaxed structure
  {oob}_slab_traj_{NNN}_energy_{E:.3f}_{formula}.traj               the relaxation trajectory

CIF files follow ASE's P1 writer (the layout of the reference's own tests/data/**/*.cif: cell lengths/angles, one
atom-site loop with type, label, multiplicity, fractional coordinates to 5 decimals, occupancy), so they read back with
``ase.io.read`` and with tests/golden/make_fixtures.py::load_cif.  ASE's ``.traj`` is its private binary ULM container:
with ASE importable the frames go through ``ase.io.trajectory.TrajectoryWriter`` exactly like the reference; without it
the same frames are written as extended-XYZ text under the same file name + ``.extxyz`` (stated divergence).
"""
from __future__ import annotations

from collections import Counter
from pathlib import Path

import numpy as np

from .engine import SYMBOLS


def _cellpar(cell):
    cell = np.asarray(cell, dtype=float).reshape(3, 3)
    lengths = np.linalg.norm(cell, axis=1)
    ang = []
    for i, j in ((1, 2), (0, 2), (0, 1)):
        c = np.dot(cell[i], cell[j]) / (lengths[i] * lengths[j]) if lengths[i] * lengths[j] > 0 else 0.0
        ang.append(np.degrees(np.arccos(np.clip(c, -1.0, 1.0))))
    return lengths, ang


def formula_hill(symbols) -> str:
    """``Atoms.get_chemical_formula()`` (Hill order; alphabetical when there is no carbon)."""
    c = Counter(symbols)
    keys = sorted(c)
    if "C" in c:
        keys = ["C"] + (["H"] if "H" in c else []) + [k for k in keys if k not in ("C", "H")]
    return "".join(f"{k}{c[k] if c[k] > 1 else ''}" for k in keys)


def _g(x):
    return f"{x:.6g}"


def write_cif(path, numbers, positions, cell) -> None:
    symbols = [SYMBOLS[int(z)] for z in numbers]
    cell = np.asarray(cell, dtype=float).reshape(3, 3)
    frac = np.asarray(positions, dtype=float) @ np.linalg.inv(cell)
    (a, b, c), (al, be, ga) = _cellpar(cell)
    cnt = Counter(symbols)
    # ASE's "reduce" formula style: runs of equal symbols in atom order
    runs, prev, n = [], None, 0
    for s in symbols + [None]:
        if s == prev:
            n += 1
        else:
            if prev is not None:
                runs.append(f"{prev}{n if n > 1 else ''}")
            prev, n = s, 1
    lines = ["data_image0",
             f"_chemical_formula_structural       {''.join(runs)}",
             "_chemical_formula_sum              \"" + " ".join(f"{k}{cnt[k]}" for k in sorted(cnt)) + "\"",
             f"_cell_length_a       {_g(a)}", f"_cell_length_b       {_g(b)}", f"_cell_length_c       {_g(c)}",
             f"_cell_angle_alpha    {_g(al)}", f"_cell_angle_beta     {_g(be)}", f"_cell_angle_gamma    {_g(ga)}",
             "", "_space_group_name_H-M_alt    \"P 1\"", "_space_group_IT_number       1", "",
             "loop_", "  _space_group_symop_operation_xyz", "  'x, y, z'", "",
             "loop_", "  _atom_site_type_symbol", "  _atom_site_label", "  _atom_site_symmetry_multiplicity",
             "  _atom_site_fract_x", "  _atom_site_fract_y", "  _atom_site_fract_z", "  _atom_site_occupancy"]
    seen = Counter()
    for s, f in zip(symbols, frac):
        seen[s] += 1
        label = f"{s}{seen[s]}"
        lines.append(f"  {s:<3s} {label:<9s} 1.0  {f[0]:.5f}  {f[1]:.5f}  {f[2]:.5f}  1.0000")
    Path(path).write_text("\n".join(lines) + "\n")


def write_traj(path, frames, cell, pbc=(True, True, True)) -> str:
    """frames: list of (numbers, positions[, energy]).  Returns the path actually written."""
    try:
        from ase import Atoms as AseAtoms
        from ase.io.trajectory import TrajectoryWriter
        w = TrajectoryWriter(str(path), mode="a")
        for fr in frames:
            w.write(AseAtoms(numbers=fr[0], positions=fr[1], cell=cell, pbc=pbc))
        w.close()
        return str(path)
    except ImportError:
        pass
    path = str(path) + ".extxyz"
    cell = np.asarray(cell, dtype=float).reshape(-1)
    with open(path, "a") as fh:
        for fr in frames:
            numbers, pos = fr[0], np.asarray(fr[1], dtype=float)
            e = f" energy={fr[2]:.8f}" if len(fr) > 2 and fr[2] is not None else ""
            fh.write(f"{len(numbers)}\n")
            fh.write('Lattice="' + " ".join(f"{x:.8f}" for x in cell) + '" Properties=species:S:1:pos:R:3' + e
                     + ' pbc="' + " ".join("T" if p else "F" for p in pbc) + '"\n')
            for z, p in zip(numbers, pos):
                fh.write(f"{SYMBOLS[int(z)]:<2s} {p[0]:16.8f} {p[1]:16.8f} {p[2]:16.8f}\n")
    return path


def save_structures(save_folder, sweep_num, energy, numbers, unrelaxed_pos, cell, relaxed_pos=None, energy_oob=False,
                    traj_frames=None, pbc=(True, True, True)) -> list[str]:
    """SurfaceSystem.save_structures (mcmc/system.py:488-534): file names and contents as in the reference."""
    folder = Path(save_folder)
    folder.mkdir(parents=True, exist_ok=True)
    formula = formula_hill([SYMBOLS[int(z)] for z in numbers])
    oob = "oob" if energy_oob else "inb"
    tag = f"{sweep_num:03}_energy_{float(energy):.3f}_{formula}"
    written = [str(folder / f"{oob}_unrelaxed_slab_sweep_{tag}.cif")]
    write_cif(written[0], numbers, unrelaxed_pos, cell)
    if relaxed_pos is not None:
        written.append(str(folder / f"{oob}_relaxed_slab_sweep_{tag}.cif"))
        write_cif(written[-1], numbers, relaxed_pos, cell)
    if traj_frames:
        written.append(write_traj(folder / f"{oob}_slab_traj_{tag}.traj", traj_frames, cell, pbc))
    return written
