"""Drop-in calculators: the reference's ASE-``Calculator`` surface over the B200 engines.

Mirrors mcmc/calculators/calculators.py of the reference (same class names, ``results`` keys,
``set``/``parameters`` config channel, ``implemented_properties`` containing ``surface_energy`` —
SURVEY.md 8b):

  EnsembleNFFSurface  (calculators.py:366-489)  PaiNN ensemble + mu/bulk-offset grand potential
  NFFPourbaix         (calculators.py:138-361)  single model + Pourbaix grand potential
  LAMMPSSurfCalc      (calculators.py:492-752)  Tersoff / SW through ``run_lammps_opt`` / ``_energy``
  get_results_single / get_std_devs_single / get_embeddings_single (calculators.py:34-135)

A single ``Atoms`` call is a batch of one.  Everything numeric runs on the GPU through the C ABI;
only the scalar thermodynamic bookkeeping (element counts x tabulated constants) is host fp64,
exactly as in the reference.
"""
from __future__ import annotations

import logging
from collections import Counter

import numpy as np

from . import engine as eng
from .atoms import Atoms, as_arrays

ENERGY_THRESHOLD = 1000  # eV
MAX_FORCE_THRESHOLD = 1000  # eV/Angstrom
ALL_CHANGES = ["positions", "numbers", "cell", "pbc", "initial_charges", "initial_magmoms"]


class PropertyNotImplementedError(NotImplementedError):
    pass


class _Parameters(dict):
    def copy(self):
        return _Parameters(self)


class Calculator:
    """The slice of ``ase.calculators.calculator.Calculator`` the reference relies on."""
    implemented_properties: tuple = ()
    name = "vssr_b200"

    def __init__(self, *args, **kwargs):
        self.parameters = _Parameters()
        self.results = {}
        self.atoms = None
        self._cache_key = None

    def set(self, **kwargs) -> dict:
        changed = {}
        for k, v in kwargs.items():
            old = self.parameters.get(k, None)
            if k not in self.parameters or not _equal(old, v):
                changed[k] = v
                self.parameters[k] = v
        if changed:
            self.reset()
        return changed

    def reset(self):
        self.results = {}
        self._cache_key = None

    def _key(self, atoms):
        pos, num, cell, pbc, _ = as_arrays(atoms)
        return (pos.tobytes(), num.tobytes(), cell.tobytes(), pbc.tobytes())

    def check_state(self, atoms) -> bool:
        """True when positions/numbers/cell/pbc changed since the cached calculation."""
        return self._cache_key is None or self._cache_key != self._key(atoms)

    def get_property(self, name, atoms=None, allow_calculation=True, **kw):
        if name not in self.implemented_properties:
            raise PropertyNotImplementedError(f"{name} property not implemented")
        if atoms is None:
            atoms = self.atoms
        changed = self.check_state(atoms)
        if changed:
            # ASE: a changed structure invalidates EVERY cached property (Calculator.get_property -> self.reset()),
            # otherwise a stale surface_energy / energy_std of the previous structure would be served
            self.results = {}
            self._cache_key = None
        if changed or name not in self.results:
            if not allow_calculation:
                return None
            self.calculate(atoms, [name], ALL_CHANGES)
        if name not in self.results:
            raise PropertyNotImplementedError(f"{name} not present in this calculation")
        r = self.results[name]
        return r.copy() if isinstance(r, np.ndarray) else r

    def get_potential_energy(self, atoms=None, **kw):
        return self.get_property("energy", atoms)

    def get_forces(self, atoms=None):
        return self.get_property("forces", atoms)

    def calculate(self, atoms=None, properties=("energy",), system_changes=ALL_CHANGES):
        if atoms is not None:
            key = self._key(atoms)
            if key != self._cache_key:
                self.results = {}      # nothing computed for another structure survives a direct calculate() either
            self.atoms = atoms.copy() if hasattr(atoms, "copy") else atoms
            self._cache_key = key


def _equal(a, b):
    try:
        r = a == b
        return bool(r) if not isinstance(r, np.ndarray) else bool(r.all())
    except Exception:
        return False


def _one_batch(atoms, type_table=None):
    pos, num, cell, pbc, fixed = as_arrays(atoms)
    z = num if type_table is None else np.array([type_table[int(q)] for q in num], dtype=np.int32)
    return eng.Batch.from_arrays([pos], [z], [cell], [pbc], [fixed]), num


# ----------------------------------------------------------------------------------------------
class EnsembleNFF(Calculator):
    """nff.io.ase_calcs.EnsembleNFF stand-in: mean/std over models."""
    implemented_properties = ("energy", "forces", "stress", "energy_std", "forces_std", "embedding")

    def __init__(self, models, device="cuda", model_units="kcal/mol", prediction_units="eV", offset_data=None,
                 cutoff=5.0, cutoff_skin=1.0, **kwargs):
        super().__init__()
        # `models`: what scripts/sample_surface.py:164-175 passes -- loaded PaiNN modules -- or state dicts
        # (checkpoint keys) or `best_model` paths; all end up as validated state dicts
        from .loaders import as_state_dict
        self.models = [as_state_dict(m) for m in models]
        self.device = device
        self.model_units, self.prediction_units = model_units, prediction_units
        self._stoich = offset_data
        self.cutoff, self.cutoff_skin = cutoff, cutoff_skin
        self._engine = None

    @property
    def engine(self) -> eng.PainnEngine:
        if self._engine is None:
            self._engine = eng.PainnEngine(self.models, self._stoich, cutoff=self.cutoff, skin=self.cutoff_skin,
                                           device=self.device)
        return self._engine

    def __deepcopy__(self, memo):
        # SurfaceSystem.copy(copy_calc=True) deep-copies the calculator (mcmc/system.py:582-584):
        # weights are immutable and shared, results/parameters are copied.
        import copy
        new = self.__class__.__new__(self.__class__)
        for k, v in self.__dict__.items():
            new.__dict__[k] = v if k in ("models", "_engine", "logger") else copy.deepcopy(v, memo)
        return new

    def calculate(self, atoms=None, properties=("energy", "forces"), system_changes=ALL_CHANGES):
        if atoms is None:
            atoms = self.atoms
        Calculator.calculate(self, atoms, properties, system_changes)
        batch, num = _one_batch(atoms)
        r = self.engine.energy_forces(batch, z_host=num, want_embedding="embedding" in properties)
        self.results = {
            "energy": r["energy"].cpu().numpy().astype(np.float32).reshape(-1),          # shape (1,)
            "energy_std": r["energy_std"].cpu().numpy().astype(np.float32).reshape(-1),
            "forces": r["forces"].cpu().numpy(),
            "forces_std": r["forces_std"].cpu().numpy(),
        }
        if r["embedding"] is not None:
            self.results["embedding"] = r["embedding"].cpu().numpy().mean(0)          # [N,128]
        if hasattr(atoms, "results"):
            atoms.results.update(self.results)


class EnsembleNFFSurface(EnsembleNFF):
    """Based on Ensemble Neural Force Field class to calculate surface energy
    (reference mcmc/calculators/calculators.py:366-489)."""
    implemented_properties = (*EnsembleNFF.implemented_properties, "surface_energy")

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.chem_pots = {}
        self.offset_data = {}
        self.offset_units = kwargs.get("offset_units", "atomic")
        self.logger = kwargs.get("logger", logging.getLogger(__name__))

    def get_surface_energy(self, atoms=None, chem_pots=None, offset_data=None) -> float:
        """Omega = E - bulk reference - sum (n_el - s_el/s_ref n_ref) mu_el (calculators.py:379-446).
        The reference's inverted if/else (ValueError when chem_pots/offset_data are passed
        explicitly) is a known oddity and is NOT replicated."""
        if atoms is None:
            atoms = self.atoms
        chem_pots = self.chem_pots if chem_pots is None else chem_pots
        offset_data = self.offset_data if offset_data is None else offset_data
        if not chem_pots:
            raise ValueError("chemical potentials are not set")
        if not offset_data:
            raise ValueError("offset data is not set")
        energy = self.get_potential_energy(atoms=atoms)
        return surface_energy_from(energy, atoms.get_chemical_symbols(), offset_data, chem_pots, self.offset_units)

    def set(self, **kwargs) -> dict:
        changed = EnsembleNFF.set(self, **kwargs)
        if "chem_pots" in self.parameters:
            self.chem_pots = self.parameters["chem_pots"]
        if "offset_data" in self.parameters:
            self.offset_data = self.parameters["offset_data"]
            if self._stoich is None and isinstance(self.offset_data, dict) and "stoidict" in self.offset_data:
                self._stoich = self.offset_data
                self._engine = None
        return changed

    def calculate(self, atoms=None, properties=implemented_properties, system_changes=ALL_CHANGES):
        if atoms is None:
            atoms = self.atoms
        EnsembleNFF.calculate(self, atoms, properties, system_changes)
        if "surface_energy" in properties:
            self.results["surface_energy"] = self.get_surface_energy(atoms=atoms)
        if hasattr(atoms, "results"):
            atoms.results.update(self.results)


def surface_energy_from(energy, symbols, offset_data, chem_pots, offset_units="atomic"):
    """H6 scalar formula (calculators.py:411-446); also used batched by the MC driver."""
    cnt = Counter(symbols)
    bulk = offset_data["bulk_energies"]
    stoics = offset_data["stoics"]
    ref_formula, ref_el = offset_data["ref_formula"], offset_data["ref_element"]
    bulk_ref = cnt[ref_el] * bulk[ref_formula]
    for el in cnt:
        if el != ref_el:
            bulk_ref += (cnt[el] - stoics[el] / stoics[ref_el] * cnt[ref_el]) * bulk[el]
    se = energy - (bulk_ref * eng.HARTREE_TO_EV if offset_units == "atomic" else bulk_ref)
    pot = 0
    for el in cnt:
        if el != ref_el:
            pot += (cnt[el] - stoics[el] / stoics[ref_el] * cnt[ref_el]) * chem_pots[el]
    return se - pot


# ----------------------------------------------------------------------------------------------
class PourbaixAtom:
    """Record used by NFFPourbaix (reference mcmc/pourbaix/atoms.py): tabulated per-element data."""

    def __init__(self, symbol, dominant_species="", species_conc=1e-6, num_e=0, num_H=0, atom_std_state_energy=0.0,
                 delta_G2_std=0.0):
        self.symbol, self.dominant_species, self.species_conc = symbol, dominant_species, species_conc
        self.num_e, self.num_H = num_e, num_H
        self.atom_std_state_energy, self.delta_G2_std = atom_std_state_energy, delta_G2_std


class NFFPourbaix(EnsembleNFF):
    """Pourbaix grand potential on top of the force field (reference calculators.py:138-361).
    Omega = -(dG1 + dG2);  dG1 = sum n_el E_std(el) - (E_slab + adsorbate corrections);
    dG2 = sum_atoms [dG2_std - n_e phi - ln10 n_H kT pH + kT ln c]."""
    implemented_properties = (*EnsembleNFF.implemented_properties, "pourbaix_potential", "surface_energy")

    def __init__(self, models, *args, **kwargs):
        models = models if isinstance(models, (list, tuple)) else [models]
        super().__init__(models, *args, **kwargs)
        self.chem_pots = {}
        self.reference_slab = {}
        self.temp = kwargs.get("temp", 0.0257)
        self.phi = kwargs.get("phi", 0)
        self.pH = kwargs.get("pH", 7)
        self.pourbaix_atoms = {}
        self.adsorbate_corrections = {}
        self.logger = kwargs.get("logger", logging.getLogger(__name__))

    def get_delta_G2_individual(self, atom) -> float:
        if isinstance(atom, str):
            atom = self.pourbaix_atoms[atom]
        non_std = (-atom.num_e * self.phi - np.log(10) * atom.num_H * self.temp * self.pH
                   + self.temp * np.log(atom.species_conc))
        return atom.delta_G2_std + non_std

    def get_delta_G2(self, atoms=None) -> float:
        atoms = self.atoms if atoms is None else atoms
        dg2 = 0
        for sym in atoms.get_chemical_symbols():
            dg2 += self.get_delta_G2_individual(sym)
        return dg2

    def get_delta_G1(self, atoms=None, slab_energy=None) -> float:
        atoms = self.atoms if atoms is None else atoms
        cnt = Counter(atoms.get_chemical_symbols())
        sum_chem = 0
        for sym, c in cnt.items():
            sum_chem += c * self.pourbaix_atoms[sym].atom_std_state_energy
        if slab_energy is None:
            slab_energy = self.get_potential_energy(atoms=atoms)
        slab_energy = slab_energy + adsorbate_correction(dict(cnt), self.adsorbate_corrections)
        return sum_chem - slab_energy

    def get_surface_energy(self, atoms=None) -> float:
        return self.get_pourbaix_potential(atoms=self.atoms if atoms is None else atoms)

    def get_pourbaix_potential(self, atoms=None) -> float:
        atoms = self.atoms if atoms is None else atoms
        return -(self.get_delta_G1(atoms=atoms) + self.get_delta_G2(atoms=atoms))

    def surface_energy_fn(self, phi=None, pH=None):
        """Host scalar for the batched MC driver: (slab energy, symbols) -> grand potential at this calculator's table,
        temperature and corrections, at the given (phi, pH) grid point (default: the calculator's own)."""
        phi = self.phi if phi is None else phi
        pH = self.pH if pH is None else pH
        table, temp, corr = self.pourbaix_atoms, self.temp, self.adsorbate_corrections or {}
        # pourbaix_potential_from with the per-species addends of this grid point formed once: the per-atom loop below
        # adds, atom by atom and in the same order, exactly the floats that function would recompute for every atom
        term = {sym: a.delta_G2_std + (-a.num_e * phi - np.log(10) * a.num_H * temp * pH + temp * np.log(a.species_conc))
                for sym, a in table.items()}
        std = {sym: a.atom_std_state_energy for sym, a in table.items()}

        def fn(e, symbols):
            cnt = Counter(symbols)
            sum_std = 0
            for sym, c in cnt.items():
                sum_std += c * std[sym]
            dg1 = sum_std - (e + adsorbate_correction(dict(cnt), corr))
            dg2 = 0
            for sym in symbols:
                dg2 += term[sym]
            return -(dg1 + dg2)

        return fn

    def set(self, **kwargs) -> dict:
        changed = EnsembleNFF.set(self, **kwargs)
        p = self.parameters
        if "temperature" in p:
            self.temp = p["temperature"]
        if "phi" in p:
            self.phi = p["phi"]
        if "pH" in p:
            self.pH = p["pH"]
        if "pourbaix_atoms" in p:
            self.pourbaix_atoms = p["pourbaix_atoms"]
        if "adsorbate_corrections" in p:
            self.adsorbate_corrections = p["adsorbate_corrections"]
        return changed

    def calculate(self, atoms=None, properties=implemented_properties, system_changes=ALL_CHANGES):
        atoms = self.atoms if atoms is None else atoms
        EnsembleNFF.calculate(self, atoms, properties, system_changes)
        if "surface_energy" in properties or "pourbaix_potential" in properties:
            v = self.get_pourbaix_potential(atoms=atoms)
            self.results["surface_energy"] = v
            self.results["pourbaix_potential"] = v
        if hasattr(atoms, "results"):
            atoms.results.update(self.results)


def pourbaix_potential_from(energy, symbols, pourbaix_atoms: dict, phi: float, pH: float, temp: float,
                            adsorbate_corrections: dict | None = None) -> float:
    """H7 scalar (calculators.py:197-305) on plain inputs: the grand potential of one structure whose slab energy is
    already known.  NFFPourbaix.get_pourbaix_potential evaluates exactly this; the multi-chain driver calls it once per
    chain with that chain's (pH, phi) grid point (BASELINE config 5)."""
    cnt = Counter(symbols)
    sum_std = 0
    for sym, c in cnt.items():
        sum_std += c * pourbaix_atoms[sym].atom_std_state_energy
    slab_energy = energy + adsorbate_correction(dict(cnt), adsorbate_corrections or {})
    dg1 = sum_std - slab_energy
    dg2 = 0
    for sym in symbols:
        a = pourbaix_atoms[sym]
        dg2 += a.delta_G2_std + (-a.num_e * phi - np.log(10) * a.num_H * temp * pH + temp * np.log(a.species_conc))
    return -(dg1 + dg2)


def adsorbate_correction(counts: dict, corrections: dict) -> float:
    """The ``Formula`` arithmetic of get_delta_G1 (calculators.py:249-272) on plain element counts:
    for every adsorbate formula, subtract excess-H water first (if it contains O and H), then add
    floor(formula / adsorbate) * correction."""
    total = 0.0
    formula = dict(counts)
    for ads, corr in corrections.items():
        ads_cnt = Counter(_formula_counts(ads))
        if "O" in ads_cnt and "H" in ads_cnt:
            ho_diff = max(formula.get("H", 0) - formula.get("O", 0), 0)
            if ho_diff > 0:
                sub = {"H": 2 * ho_diff, "O": ho_diff}
                formula = {k: v - sub.get(k, 0) for k, v in formula.items()}
        div = min((formula.get(k, 0) // v for k, v in ads_cnt.items()), default=0)
        div = max(div, 0)
        total += div * corr
    return total


def _formula_counts(s: str):
    import re
    out = []
    for sym, c in re.findall(r"([A-Z][a-z]?)(\d*)", s):
        out += [sym] * (int(c) if c else 1)
    return out


# ----------------------------------------------------------------------------------------------
class LAMMMPSCalc(Calculator):
    """The reference's in-process LAMMPS calculator (calculators.py:492-693), served by the
    Tersoff / SW CUDA kernels instead of template files + a new LAMMPS instance per call."""
    name = "lammpscalc"
    implemented_properties = ("energy", "relaxed_energy", "forces", "per_atom_energies")

    def __init__(self, potential="tersoff", params=None, elements=("Ga", "N"), bulk_index=0, n_max=128, max_nbr=32,
                 device="cuda", **kwargs):
        super().__init__()
        self.run_dir = "."
        self.relax_steps = 100
        self.kim_potential = False
        self.logger = kwargs.get("logger", logging.getLogger(__name__))
        self.elements = list(elements)
        self.bulk_index = bulk_index          # `group bulk id <= bulk_index` + `fix setforce 0`
        self.type_table = {eng.NUMBERS[e]: k for k, e in enumerate(self.elements)}
        kind = eng.POT_TERSOFF if potential == "tersoff" else eng.POT_SW
        self._engine_args = (kind, np.asarray(params, dtype=np.float64), len(self.elements), n_max, max_nbr)
        self._device = device
        self._engine = None

    @property
    def engine(self) -> eng.ClassicalEngine:
        if self._engine is None:
            k, p, nt, nm, mn = self._engine_args
            self._engine = eng.ClassicalEngine(k, p, nt, n_max=nm, max_nbr=mn, device=self._device)
        return self._engine

    def _batch(self, slab):
        pos, num, cell, pbc, _ = as_arrays(slab)
        types = np.array([self.type_table[int(q)] for q in num], dtype=np.int32)
        fixed = np.arange(len(num)) < self.bulk_index
        return eng.Batch.from_arrays([pos], [types], [cell], [pbc], [fixed])

    def run_lammps_opt(self, slab, run_dir="./", **kwargs):
        """-> (opt_slab, energy, pe_per_atom) like calculators.py:600-619 (minimise with bulk frozen)."""
        steps = kwargs.get("relax_steps", self.relax_steps)
        b = self._batch(slab)
        r = self.engine.relax(b, relax_steps=steps, fmax=kwargs.get("fmax", 0.01))
        out = r["out"].cpu().numpy()[0]
        opt = slab.copy()
        opt.positions = b.pos.cpu().numpy() if isinstance(opt, Atoms) else opt.positions
        if not isinstance(opt, Atoms):
            opt.set_positions(b.pos.cpu().numpy(), apply_constraint=False)
        opt.calc = getattr(slab, "calc", None)
        self._last_forces = r["forces"].cpu().numpy()
        return opt, float(out[2]), []

    def run_lammps_energy(self, slab, run_dir="./", **kwargs):
        b = self._batch(slab)
        r = self.engine.energy_forces(b)
        self._last_forces = r["forces"].cpu().numpy()
        return slab, float(r["energy"].item()), r["per_atom_energies"].cpu().numpy()

    def set(self, **kwargs) -> dict:
        changed = Calculator.set(self, **kwargs)
        if "run_dir" in self.parameters:
            self.run_dir = self.parameters["run_dir"]
        if "relax_steps" in self.parameters:
            self.relax_steps = self.parameters["relax_steps"]
        if "kim_potential" in self.parameters:
            self.kim_potential = self.parameters["kim_potential"]
        return changed

    def calculate(self, atoms=None, properties=implemented_properties, system_changes=ALL_CHANGES):
        atoms = self.atoms if atoms is None else atoms
        Calculator.calculate(self, atoms, properties, system_changes)
        if "energy" in properties or "forces" in properties or "per_atom_energies" in properties:
            _, e, pe = self.run_lammps_energy(atoms, run_dir=self.run_dir)
            self.results["energy"] = e
            self.results["per_atom_energies"] = pe
            self.results["forces"] = self._last_forces
        if "relaxed_energy" in properties:
            _, e, pe = self.run_lammps_opt(atoms, run_dir=self.run_dir)
            self.results["relaxed_energy"] = e


class LAMMPSSurfCalc(LAMMMPSCalc):
    """Surface energy = potential energy (calculators.py:696-752)."""
    implemented_properties = (*LAMMMPSCalc.implemented_properties, "surface_energy")

    def get_surface_energy(self, atoms=None) -> float:
        """calculators.py:712-727: currently the same as the potential energy (served from the cache when `atoms` is
        the structure the cached results belong to, like ASE's get_potential_energy)."""
        return self.get_potential_energy(atoms=self.atoms if atoms is None else atoms)

    def calculate(self, atoms=None, properties=implemented_properties, system_changes=ALL_CHANGES):
        atoms = self.atoms if atoms is None else atoms
        LAMMMPSCalc.calculate(self, atoms, properties, system_changes)
        if "surface_energy" in properties:
            if "energy" not in self.results:       # results of ANOTHER structure were dropped by Calculator.calculate
                LAMMMPSCalc.calculate(self, atoms, ("energy",), system_changes)
            self.results["surface_energy"] = self.results["energy"]


class LAMMPSRunSurfCalc(Calculator):
    """``LAMMPSRunSurfCalc`` (calculators.py:755-811): ASE's file-based LAMMPS runner with ``surface_energy`` =
    potential energy, as used by the Cu(100) / Au(110) toy runs (tutorials/example.ipynb, tests/test_Cu.py,
    tests/test_Au.py) with ``pair_style eam`` and a single-element funcfl file.  Served by the EAM branch of the
    classical CUDA kernel instead of an ``lmp`` subprocess per call; no relaxation on this path (the reference runs it
    with ``relax_atoms`` unset)."""
    name = "lammpsrun"
    implemented_properties = ("energy", "free_energy", "forces", "energies", "surface_energy")

    def __init__(self, files=None, funcfl=None, n_max=32, max_nbr=128, device="cuda", **kwargs):
        super().__init__()
        self.files = [str(f) for f in (files or [])]
        self._funcfl = funcfl                 # parsed table (loaders.load_eam_funcfl) or None -> read from `files`
        self._engine_cfg = (n_max, max_nbr, device)
        self._engine = None
        self.run_dir = kwargs.get("tmp_dir", ".")
        self.logger = kwargs.get("logger", logging.getLogger(__name__))

    @property
    def engine(self) -> eng.ClassicalEngine:
        if self._engine is None:
            style = self.parameters.get("pair_style", "eam")
            if style != "eam":
                raise NotImplementedError(f"pair_style {style!r}: LAMMPSRunSurfCalc serves `eam` (funcfl) on the B200 engine")
            tab = self._funcfl
            if tab is None:
                from .loaders import load_eam_funcfl
                coeff = self.parameters.get("pair_coeff", ["* * "])[0].split()[-1]      # "* * Cu_u3.eam"
                match = [f for f in self.files if f.endswith(coeff)] or self.files
                if not match:
                    raise FileNotFoundError("no EAM potential file: pass files=[...] or funcfl=")
                tab = load_eam_funcfl(match[0])
            n_max, max_nbr, device = self._engine_cfg
            self._engine = eng.ClassicalEngine(eng.POT_EAM, eng.eam_param_block(tab), 1, n_max=n_max, max_nbr=max_nbr,
                                               device=device)
        return self._engine

    def get_surface_energy(self, atoms=None) -> float:
        return self.get_potential_energy(atoms=self.atoms if atoms is None else atoms)

    def calculate(self, atoms=None, properties=implemented_properties, system_changes=ALL_CHANGES):
        atoms = self.atoms if atoms is None else atoms
        Calculator.calculate(self, atoms, properties, system_changes)
        pos, num, cell, pbc, _ = as_arrays(atoms)
        if len(set(num.tolist())) > 1:
            raise NotImplementedError("funcfl EAM is a single-element potential")
        b = eng.Batch.from_arrays([pos], [np.zeros(len(num), np.int32)], [cell], [pbc])
        if len(num) > self._engine_cfg[0]:          # structure outgrew the shared-memory budget: re-size the engine
            self._engine_cfg = (int(len(num) * 1.25) + 8, *self._engine_cfg[1:])
            self._engine = None
        r = self.engine.energy_forces(b)
        e = float(r["energy"].item())
        self.results.update({"energy": e, "free_energy": e, "forces": r["forces"].cpu().numpy(),
                             "energies": r["per_atom_energies"].cpu().numpy()})
        if "surface_energy" in properties:
            self.results["surface_energy"] = e


# ----------------------------------------------------------------------------------------------
def get_results_single(atoms_batch, calc) -> dict:
    """calculators.py:34-47"""
    atoms_batch.calc = calc
    calc.calculate(atoms_batch)
    return calc.results


def get_embeddings_single(atoms_batch, calc, results_cache=None, flatten=True, flatten_axis=0) -> np.ndarray:
    """calculators.py:68-95"""
    if results_cache is not None and "embedding" in results_cache:
        results = results_cache
    else:
        atoms_batch.calc = calc
        calc.calculate(atoms_batch, properties=("energy", "forces", "embedding"))
        results = calc.results
    e = results["embedding"]
    return e.mean(axis=flatten_axis).squeeze() if flatten else e.squeeze()


def get_std_devs_single(atoms_batch, calc) -> np.ndarray:
    """calculators.py:118-135"""
    if len(calc.models) > 1:
        atoms_batch.calc = calc
        calc.calculate(atoms_batch)
        return calc.results.get("forces_std", np.array([0.0])).mean()
    return 0.0
