"""Minimal ``ase.Atoms`` stand-in for the calculator boundary.

The reference hands ``ase.Atoms`` / NFF ``AtomsBatch`` objects to its calculators
(mcmc/system.py:184-234, mcmc/dynamics.py:83-170).  ASE is not installable in this image, so the
boundary accepts *either* a real ``ase.Atoms`` (duck-typed: ``get_positions``,
``get_atomic_numbers``, ``get_cell``, ``get_pbc``, ``constraints``) or this shim, which implements
just the members the hot path touches, with ASE's index semantics (append at end, ``del`` shifts).
"""
from __future__ import annotations

import numpy as np

from .engine import NUMBERS, SYMBOLS


class FixAtoms:
    def __init__(self, indices=None, mask=None):
        if mask is not None:
            indices = np.where(np.asarray(mask, dtype=bool))[0]
        self.index = np.asarray(indices if indices is not None else [], dtype=int)

    def get_indices(self):
        return self.index

    def todict(self):
        return {"name": "FixAtoms", "kwargs": {"indices": self.index.tolist()}}


class Atoms:
    def __init__(self, symbols=None, positions=None, numbers=None, cell=None, pbc=False, constraint=None,
                 calculator=None):
        if numbers is None:
            numbers = [NUMBERS[s] for s in _parse_symbols(symbols)] if symbols is not None else []
        self.numbers = np.asarray(numbers, dtype=int).copy()
        n = len(self.numbers)
        self.positions = (np.zeros((n, 3)) if positions is None else np.asarray(positions, dtype=float).reshape(n, 3).copy())
        self.cell = np.zeros((3, 3)) if cell is None else np.asarray(cell, dtype=float).reshape(3, 3).copy()
        self.pbc = np.broadcast_to(np.asarray(pbc, dtype=bool), (3,)).copy()
        self.constraints = []
        if constraint is not None:
            self.set_constraint(constraint)
        self.calc = calculator
        self.arrays = {}
        self.results = {}
        self.info = {}

    # --- ASE-like API -----------------------------------------------------------------
    def __len__(self):
        return len(self.numbers)

    def copy(self):
        a = Atoms(numbers=self.numbers, positions=self.positions, cell=self.cell, pbc=self.pbc)
        a.constraints = [FixAtoms(indices=c.index.copy()) for c in self.constraints]
        a.arrays = {k: v.copy() for k, v in self.arrays.items()}
        a.info = dict(self.info)
        return a

    def get_positions(self, wrap=False):
        return self.positions.copy()

    def set_positions(self, p):
        p = np.asarray(p, dtype=float)
        fixed = self.fixed_mask()
        p = p.copy()
        p[fixed] = self.positions[fixed]
        self.positions = p

    def get_atomic_numbers(self):
        return self.numbers.copy()

    def get_chemical_symbols(self):
        return [SYMBOLS[int(z)] for z in self.numbers]

    def get_chemical_formula(self):
        from collections import Counter
        c = Counter(self.get_chemical_symbols())
        return "".join(f"{k}{v if v > 1 else ''}" for k, v in sorted(c.items()))

    def get_cell(self):
        return self.cell.copy()

    def get_pbc(self):
        return self.pbc.copy()

    def set_constraint(self, constraint=None):
        if constraint is None:
            self.constraints = []
        elif isinstance(constraint, (list, tuple)):
            self.constraints = list(constraint)
        else:
            self.constraints = [constraint]

    def fixed_mask(self) -> np.ndarray:
        m = np.zeros(len(self), dtype=bool)
        for c in self.constraints:
            idx = np.asarray(c.index, dtype=int)
            m[idx[idx < len(self)]] = True
        return m

    def set_array(self, name, a, dtype=None):
        self.arrays[name] = np.asarray(a, dtype=dtype).copy()

    def get_array(self, name):
        return self.arrays[name].copy()

    def append(self, symbol_or_z, position=None):
        z = NUMBERS[symbol_or_z] if isinstance(symbol_or_z, str) else int(symbol_or_z)
        self.numbers = np.append(self.numbers, z)
        self.positions = np.vstack([self.positions, np.zeros(3) if position is None else np.asarray(position, float)])
        for k, v in self.arrays.items():
            self.arrays[k] = np.append(v, np.zeros(1, dtype=v.dtype))

    def __delitem__(self, i):
        keep = np.ones(len(self), dtype=bool)
        keep[i] = False
        new_index = np.cumsum(keep) - 1
        self.numbers = self.numbers[keep]
        self.positions = self.positions[keep]
        for k, v in self.arrays.items():
            self.arrays[k] = v[keep]
        for c in self.constraints:
            idx = c.index[keep[c.index]] if len(c.index) else c.index
            c.index = new_index[idx]

    def get_potential_energy(self, **kw):
        if self.calc is None:
            raise RuntimeError("Atoms object has no calculator.")
        return self.calc.get_potential_energy(atoms=self)

    def get_forces(self, apply_constraint=True, **kw):
        if self.calc is None:
            raise RuntimeError("Atoms object has no calculator.")
        f = np.array(self.calc.get_forces(atoms=self), dtype=float)
        if apply_constraint:
            f[self.fixed_mask()] = 0.0
        return f


def _parse_symbols(symbols):
    if isinstance(symbols, (list, tuple)):
        return list(symbols)
    import re
    out = []
    for sym, cnt in re.findall(r"([A-Z][a-z]?)(\d*)", symbols):
        out += [sym] * (int(cnt) if cnt else 1)
    return out


def as_arrays(atoms):
    """(positions[N,3] f64, numbers[N] int, cell[3,3], pbc[3], fixed_mask[N]) from ase.Atoms or the shim."""
    pos = np.asarray(atoms.get_positions(), dtype=np.float64)
    num = np.asarray(atoms.get_atomic_numbers(), dtype=np.int64)
    cell = np.asarray(atoms.get_cell(), dtype=np.float64).reshape(3, 3)
    pbc = np.asarray(atoms.get_pbc(), dtype=bool)
    fixed = np.zeros(len(num), dtype=bool)
    for c in getattr(atoms, "constraints", []) or []:
        idx = getattr(c, "index", None)
        if idx is None and hasattr(c, "get_indices"):
            idx = c.get_indices()
        if idx is not None:
            idx = np.asarray(idx, dtype=int)
            fixed[idx[idx < len(num)]] = True
    return pos, num, cell, pbc, fixed
