"""Result-cache semantics of the ASE-Calculator stand-in (no GPU: the engine calls are mocked).

Regressions for two stale-result bugs: a calculator that served structure A's surface_energy for
structure B made every Metropolis step after the first see dE = 0."""
import numpy as np

from surface_sampling_b200 import calculators as calcs
from surface_sampling_b200.atoms import Atoms


def _atoms(x):
    return Atoms(numbers=[31, 7], positions=[[0, 0, 0], [x, 0, 0]], cell=np.eye(3) * 10, pbc=True)


class _MockLammps(calcs.LAMMPSSurfCalc):
    def __init__(self):
        calcs.Calculator.__init__(self)
        self.run_dir, self.relax_steps, self.n_calls = ".", 100, 0

    def run_lammps_energy(self, slab, run_dir="./", **kw):
        self.n_calls += 1
        self._last_forces = np.zeros((len(slab), 3))
        return slab, float(slab.get_positions()[1, 0]), np.zeros(len(slab))


def test_lammps_surf_calc_recomputes_for_a_new_structure():
    c = _MockLammps()
    a, b = _atoms(1.0), _atoms(2.0)
    assert c.get_property("surface_energy", atoms=a) == 1.0
    assert c.get_property("surface_energy", atoms=b) == 2.0          # was 1.0: stale 'energy' copied over
    assert c.get_property("energy", atoms=b) == 2.0
    n = c.n_calls
    assert c.get_property("surface_energy", atoms=b) == 2.0 and c.n_calls == n   # unchanged structure: cached
    assert c.get_property("energy", atoms=a) == 1.0 and "surface_energy" not in c.results


class _MockEnsemble(calcs.EnsembleNFFSurface):
    def __init__(self):
        calcs.Calculator.__init__(self)
        self.chem_pots, self.offset_data, self.offset_units = {"N": 0.0}, None, "atomic"

    def calculate(self, atoms=None, properties=("energy",), system_changes=calcs.ALL_CHANGES):
        calcs.Calculator.calculate(self, atoms, properties, system_changes)
        e = float(atoms.get_positions()[1, 0])
        self.results = {"energy": np.array([e], np.float32), "forces": np.zeros((len(atoms), 3)),
                        "energy_std": np.array([0.5 * e], np.float32)}
        if "surface_energy" in properties:
            self.results["surface_energy"] = 100.0 + e


def test_get_property_drops_every_cached_key_on_a_changed_structure():
    c = _MockEnsemble()
    a, b = _atoms(1.0), _atoms(2.0)
    assert c.get_property("surface_energy", atoms=a) == 101.0
    assert c.get_property("energy", atoms=b)[0] == 2.0
    assert "surface_energy" not in c.results                           # A's value must not survive
    assert c.get_property("surface_energy", atoms=b) == 102.0
    assert c.get_property("energy_std", atoms=a)[0] == 0.5


def test_optimize_slab_primes_the_calculator_without_leftovers(monkeypatch):
    """optimize_slab leaves the calculator primed with the relaxed energy/forces ONLY (advice r1: results.update kept
    the previous structure's surface_energy)."""
    from surface_sampling_b200 import dynamics

    class Eng:
        cutoff, skin = 5.0, 1.0

        def relax(self, batch, relax_steps, fmax, z_host, check):
            import torch
            e = float(batch.pos[1, 0])
            return {"out": torch.tensor([[e, 0, e, 0.1, 3, 1, 0, 4]], dtype=torch.float64),
                    "forces": torch.zeros((batch.n_atoms, 3))}

    class B:
        def __init__(self, pos):
            import torch
            self.pos, self.n_atoms = torch.tensor(pos), len(pos)

    monkeypatch.setattr(dynamics.eng.Batch, "from_arrays", staticmethod(lambda p, z, c, pb, f: B(p[0])))
    c = _MockEnsemble()
    c._engine = Eng()
    for x in (1.0, 2.0):
        a = _atoms(x)
        a.calc = c
        slab, _, e, oob = dynamics.optimize_slab(a, optimizer="FIRE", save_traj=False)
        assert e == x and not oob
        assert c.get_property("surface_energy", atoms=slab) == 100.0 + x   # second pass returned 101.0 before the fix
