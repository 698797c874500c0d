"""Tersoff / Stillinger-Weber CUDA kernels (fp64) vs the CPU oracle and the GaN golden value."""
import numpy as np
import pytest
import torch

from conftest import perturbed, with_adsorbates
from oracle import classical as ocl
from oracle import relax as orelax

pytestmark = pytest.mark.gpu


def _types(numbers, table):
    return np.array([table[int(z)] for z in numbers], dtype=np.int32)


def _batch(structs, types, fixed=None):
    from surface_sampling_b200 import engine
    return engine.Batch.from_arrays([s["positions"] for s in structs], types, [s["cell"] for s in structs],
                                    [s["pbc"] for s in structs], fixed)


def test_tersoff_energy_forces(structures, potentials, golden_values):
    from surface_sampling_b200 import engine
    tab = engine.tersoff_param_table(potentials["GaN.tersoff"], ["Ga", "N"])
    eng = engine.ClassicalEngine(engine.POT_TERSOFF, tab, 2, n_max=64, max_nbr=24)
    prm = ocl.TersoffParams(potentials["GaN.tersoff"], ["Ga", "N"])
    rng = np.random.default_rng(1)
    base = structures["GaN_0001_3x3"]
    structs = [base, perturbed(base, rng, 0.08), with_adsorbates(perturbed(base, rng, 0.05), rng, 12, [31], 1.8)]
    types = [_types(s["numbers"], {31: 0, 7: 1}) for s in structs]
    b = _batch(structs, types)
    r = eng.energy_forces(b)
    e = r["energy"].cpu().numpy()
    f = b.split_host(r["forces"].cpu().numpy())
    assert abs(e[0] - golden_values["tersoff_gan_pristine"]["energy"]) < 1e-3
    for k, s in enumerate(structs):
        e0, f0 = ocl.energy_forces(ocl.tersoff_energy, s["positions"], torch.tensor(types[k]).long(), s["cell"],
                                   s["pbc"], prm)
        assert abs(e[k] - e0) < 1e-9 * max(1.0, abs(e0)), (k, e[k], e0)
        assert np.abs(f[k] - f0).max() < 1e-8, (k, np.abs(f[k] - f0).max())
    eat = b.split_host(r["per_atom_energies"].cpu().numpy())
    assert abs(eat[0].sum() - e[0]) < 1e-10


def test_sw_energy_forces(structures):
    from surface_sampling_b200 import engine
    eng = engine.ClassicalEngine(engine.POT_SW, engine.sw_param_table(), 1, n_max=128, max_nbr=32)
    rng = np.random.default_rng(4)
    base = structures["Si_111_5x5"]
    structs = [base, perturbed(base, rng, 0.1), with_adsorbates(base, rng, 5, [14], 2.0)]
    types = [np.zeros(len(s["numbers"]), np.int32) for s in structs]
    b = _batch(structs, types)
    r = eng.energy_forces(b)
    e = r["energy"].cpu().numpy()
    f = b.split_host(r["forces"].cpu().numpy())
    for k, s in enumerate(structs):
        e0, f0 = ocl.energy_forces(ocl.sw_energy, s["positions"], s["cell"], s["pbc"], ocl.SWParams())
        assert abs(e[k] - e0) < 1e-9 * abs(e0), (k, e[k], e0)
        assert np.abs(f[k] - f0).max() < 1e-8


@pytest.mark.parametrize("kind", ["tersoff", "sw"])
def test_relax_vs_oracle_fire(structures, potentials, kind):
    from surface_sampling_b200 import engine
    rng = np.random.default_rng(9)
    if kind == "tersoff":
        tab = engine.tersoff_param_table(potentials["GaN.tersoff"], ["Ga", "N"])
        eng = engine.ClassicalEngine(engine.POT_TERSOFF, tab, 2, n_max=64, max_nbr=24)
        prm = ocl.TersoffParams(potentials["GaN.tersoff"], ["Ga", "N"])
        base = structures["GaN_0001_3x3"]
        structs = [with_adsorbates(base, rng, 6, [31], 1.8), with_adsorbates(base, rng, 12, [31], 1.8)]
        types = [_types(s["numbers"], {31: 0, 7: 1}) for s in structs]
        fixed = [np.arange(len(s["numbers"])) < 36 for s in structs]   # bulk_index 36 (GaN_0001_lammps_config.json)
        steps = 100

        def efn(k):
            return lambda x: ocl.energy_forces(ocl.tersoff_energy, x, torch.tensor(types[k]).long(),
                                               structs[k]["cell"], structs[k]["pbc"], prm)
    else:
        eng = engine.ClassicalEngine(engine.POT_SW, engine.sw_param_table(), 1, n_max=128, max_nbr=32)
        base = structures["Si_111_5x5"]
        structs = [perturbed(base, rng, 0.05), with_adsorbates(base, rng, 4, [14], 2.0)]
        types = [np.zeros(len(s["numbers"]), np.int32) for s in structs]
        fixed = [np.arange(len(s["numbers"])) < 75 for s in structs]   # bulk_index 75 (Si_111_5x5_lammps_config.json)
        steps = 40

        def efn(k):
            return lambda x: ocl.energy_forces(ocl.sw_energy, x, structs[k]["cell"], structs[k]["pbc"], ocl.SWParams())
    b = _batch(structs, types, fixed)
    res = eng.relax(b, relax_steps=steps, fmax=0.01)
    out = res["out"].cpu().numpy()
    pos = b.split_host(b.pos.cpu().numpy())
    for k, s in enumerate(structs):
        o = orelax.relax(efn(k), s["positions"], fixed[k], optimizer="FIRE", relax_steps=steps, fmax=0.01)
        assert int(out[k, 4]) == o["nsteps"] and bool(out[k, 5]) == o["converged"], (out[k], o["nsteps"])
        assert abs(out[k, 2] - o["raw_energy"]) < 1e-7 * abs(o["raw_energy"]), (out[k, 2], o["raw_energy"])
        assert np.abs(pos[k] - o["pos"]).max() < 1e-6
        assert np.array_equal(pos[k][fixed[k]], s["positions"][fixed[k]])


def test_host_buffer_entry_matches_device_entry(structures, potentials):
    """vssr_classical_relax_host: the FFI-facing call with host pointers."""
    import ctypes as C
    from surface_sampling_b200 import _lib, engine
    lib = _lib.load()
    tab = np.ascontiguousarray(engine.tersoff_param_table(potentials["GaN.tersoff"], ["Ga", "N"]))
    rng = np.random.default_rng(3)
    s = with_adsorbates(structures["GaN_0001_3x3"], rng, 12, [31], 1.8)
    types = _types(s["numbers"], {31: 0, 7: 1})
    fixed = (np.arange(len(types)) < 36).astype(np.uint8)
    eng = engine.ClassicalEngine(engine.POT_TERSOFF, tab, 2, n_max=64, max_nbr=24)
    b = _batch([s], [types], [fixed])
    ref = eng.relax(b, relax_steps=30)["out"].cpu().numpy()
    pos = np.ascontiguousarray(s["positions"], dtype=np.float64).copy()
    ptr = np.array([0, len(types)], np.int32)
    cell = np.ascontiguousarray(s["cell"], dtype=np.float64)
    pbc = np.ascontiguousarray(s["pbc"]).astype(np.uint8)
    out = np.zeros((1, 8)); forces = np.zeros_like(pos); status = np.zeros(1, np.int32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = lib.vssr_classical_relax_host(0, p(tab), 2, p(pos), p(types), p(fixed), p(ptr), p(cell), p(pbc), 1,
                                       len(types), 64, 24, 30, 0.01, 0.5, p(out), p(forces), p(status))
    assert rc == 0 and status[0] == 0
    assert np.array_equal(out, ref)
    assert np.array_equal(pos, b.pos.cpu().numpy())


def _eam(el):
    from surface_sampling_b200 import engine
    from oracle.eam import EAMFuncfl
    from conftest import GOLDEN
    z = np.load(GOLDEN / "eam_funcfl.npz")
    tab = {k.split("/")[1]: z[k] for k in z.files if k.startswith(el + "/")}
    return engine.ClassicalEngine(engine.POT_EAM, engine.eam_param_block(tab), 1, n_max=32, max_nbr=128), EAMFuncfl(tab), tab


def test_eam_au_golden_and_oracle(structures, golden_values):
    """BASELINE config 1 family (LAMMPSRunSurfCalc, `pair_style eam`): the 28 canonical Au(110) states of the reference's
    tests/test_Au.py in ONE batch through the CUDA kernel: min = -79.03490823689619 (tests/test_Au.py:19), and every
    state's energy / forces / per-atom energies vs the oracle."""
    import itertools
    eng, au, _ = _eam("Au")
    slab = structures["Au_110_2x2"]
    ads = structures["Au_110_2x2_proper_adsorbed"]["positions"][16:24]
    structs = [{"positions": np.concatenate([slab["positions"], ads[list(keep)]]), "cell": slab["cell"], "pbc": slab["pbc"]}
               for keep in itertools.combinations(range(8), 6)]
    rng = np.random.default_rng(3)
    structs.append({**structs[0], "positions": structs[0]["positions"] + rng.normal(0, 0.1, structs[0]["positions"].shape)})
    b = _batch(structs, [np.zeros(len(s["positions"]), np.int32) for s in structs])
    r = eng.energy_forces(b)
    e = r["energy"].cpu().numpy()
    f = b.split_host(r["forces"].cpu().numpy())
    eat = b.split_host(r["per_atom_energies"].cpu().numpy())
    assert abs(e[:28].min() - golden_values["eam_au"]["energy"]) < 1e-5      # CIF adatoms carry 5 decimals
    for k, s in enumerate(structs):
        e0, f0 = au.energy_forces(s["positions"], s["cell"], s["pbc"])
        assert abs(e[k] - e0) < 1e-9 * abs(e0), (k, e[k], e0)
        assert np.abs(f[k] - f0).max() < 1e-8, (k, np.abs(f[k] - f0).max())
        assert abs(eat[k].sum() - e[k]) < 1e-10
    assert np.abs(f[28]).max() > 0.1                                          # the perturbed state has real forces


def test_eam_cu_calculator_and_toy_mc(golden_values):
    """BASELINE config 1: Cu(100) toy VSSR-MC through the reference-named calculator.  tests/test_Cu.py:19 asserts
    min(energy_hist) = -25.2893, which is the slab + one bridge Cu; the product calculator reproduces it, and a short
    unrelaxed MC run over {ontop, bridge} sites makes the oracle loop's decisions."""
    from test_oracle_golden import cu100_slab
    from oracle import mc as omc
    from surface_sampling_b200 import mc
    from surface_sampling_b200.atoms import Atoms
    from surface_sampling_b200.calculators import LAMMPSRunSurfCalc
    eng, cu, tab = _eam("Cu")
    g = golden_values["eam_cu"]
    pos, cell, pbc = cu100_slab(g["a"])
    d, zt = g["a"] / np.sqrt(2), pos[:, 2].max() + g["planar_distance"]
    calc = LAMMPSRunSurfCalc(funcfl=tab, keep_tmp_files=False, keep_alive=False)
    assert set(calc.set(pair_style="eam", pair_coeff=["* * Cu_u3.eam"])) == {"pair_style", "pair_coeff"}
    slab = Atoms(numbers=[29] * 8, positions=pos, cell=cell, pbc=pbc, calculator=calc)
    ads = slab.copy(); ads.append("Cu", [0.5 * d, 0.0, zt]); ads.calc = calc
    e = calc.get_property("surface_energy", atoms=ads)
    assert np.allclose(e, g["energy"]) and abs(e - g["energy"]) < 1e-4
    assert calc.get_potential_energy(atoms=slab) != e and calc.results["forces"].shape == (8, 3)   # cache follows the structure
    assert abs(calc.results["energies"].sum() - calc.results["energy"]) < 1e-10
    sites = np.array([[0, 0, zt], [d, 0, zt], [0, d, zt], [d, d, zt],
                      [0.5 * d, 0, zt], [1.5 * d, 0, zt], [0, 0.5 * d, zt], [0, 1.5 * d, zt]])

    def relax_fn(pos_l, num_l, fix_l):      # LAMMPSRunSurfCalc runs unrelaxed: energy only
        from surface_sampling_b200 import engine
        b = engine.Batch.from_arrays(pos_l, [np.zeros(len(z), np.int32) for z in num_l], [cell] * len(pos_l), [pbc] * len(pos_l))
        out = np.zeros((len(pos_l), 8))
        out[:, 2] = out[:, 0] = eng.energy_forces(b)["energy"].cpu().numpy()
        return out

    seeds = [0, 1, 2, 3]
    drv = mc.MultiChainMC([29] * 8, pos, np.ones(8, bool), sites, ["Cu"], relax_fn, lambda e_, sym: e_, seeds)
    res = drv.run(total_sweeps=10, sweep_size=2, start_temp=1.0, perform_annealing=True, alpha=0.99)   # tests/test_Cu.py:53-61
    for k, sd in enumerate(seeds):
        o = omc.run_chain(sd, ["Cu"] * 8, pos, sites, ["Cu"], lambda sym, p: cu.energy_forces(p, cell, pbc)[0], 10, 2,
                          start_temp=1.0, alpha=0.99)
        assert [x[0] for x in drv.decisions[k]] == [x[0] for x in o["decisions"]]
        assert np.allclose([x[1] for x in drv.decisions[k]], [x[1] for x in o["decisions"]], rtol=1e-10, atol=1e-9)
        assert np.allclose(res["energy_hist"][k], o["energy_hist"], rtol=1e-10, atol=1e-9)


def test_pair_table_overflow_is_reported(structures):
    """Tersoff / SW keep the directed pairs inside the potential cutoff in a shared-memory table of 8 * n_max entries.
    The Si slab compressed by 5 % pulls the second shell inside the SW cutoff (12.5 pairs per atom, 1250 in all): with
    n_max = 100 that does not fit and the kernel must say so through the status word (VssrError), never return numbers
    from a truncated table; with n_max = 160 (1280 entries) it runs and matches the oracle."""
    from surface_sampling_b200 import _lib, engine
    base = structures["Si_111_5x5"]
    dense = dict(base, positions=base["positions"] * 0.95, cell=np.asarray(base["cell"]) * 0.95)
    n = len(base["numbers"])
    assert n == 100
    types = [np.zeros(n, np.int32)]
    tight = engine.ClassicalEngine(engine.POT_SW, engine.sw_param_table(), 1, n_max=n, max_nbr=24)
    with pytest.raises(_lib.VssrError):
        tight.energy_forces(_batch([dense], types))
    roomy = engine.ClassicalEngine(engine.POT_SW, engine.sw_param_table(), 1, n_max=160, max_nbr=24)
    r = roomy.energy_forces(_batch([dense], types))
    e0, f0 = ocl.energy_forces(ocl.sw_energy, dense["positions"], dense["cell"], dense["pbc"], ocl.SWParams())
    assert abs(r["energy"].cpu().numpy()[0] - e0) < 1e-9 * abs(e0)
    assert np.abs(r["forces"].cpu().numpy() - f0).max() < 1e-7


@pytest.mark.parametrize("kind", ["tersoff", "sw"])
def test_edge_cases_isolated_atoms_and_size_limit(structures, potentials, kind):
    """Ragged batch: an isolated atom (no pair inside the cutoff), a dimer outside the cutoff, a bonded dimer and a normal
    slab in one launch -- zero energy and force where nothing interacts, the slab unchanged by its neighbours in the
    batch; a structure larger than n_max is reported, not truncated silently."""
    from surface_sampling_b200 import _lib, engine
    if kind == "tersoff":
        eng = engine.ClassicalEngine(engine.POT_TERSOFF, engine.tersoff_param_table(potentials["GaN.tersoff"], ["Ga", "N"]), 2,
                                     n_max=64, max_nbr=32)
        base, z, d_bond = structures["GaN_0001_3x3"], 31, 2.4
        tab = {31: 0, 7: 1}
    else:
        eng = engine.ClassicalEngine(engine.POT_SW, engine.sw_param_table(), 1, n_max=128, max_nbr=32)
        base, z, d_bond = structures["Si_111_5x5"], 14, 2.35
        tab = {14: 0}
    box = np.eye(3) * 30.0
    pbc = np.array([True, True, True])
    mk = lambda pts: {"positions": np.array(pts, float), "numbers": np.full(len(pts), z), "cell": box, "pbc": pbc}
    structs = [mk([[1, 1, 1]]), mk([[1, 1, 1], [9, 1, 1]]), mk([[1, 1, 1], [1 + d_bond, 1, 1]]), base]
    types = [np.array([tab[int(q)] for q in s["numbers"]], np.int32) for s in structs]
    b = _batch(structs, types)
    r = eng.energy_forces(b)
    e = r["energy"].cpu().numpy()
    f = b.split_host(r["forces"].cpu().numpy())
    assert e[0] == 0.0 and e[1] == 0.0 and not f[0].any() and not f[1].any()
    assert e[2] != 0.0 and np.abs(f[2][0] + f[2][1]).max() < 1e-12 and abs(f[2][0][0]) > 1e-3 and not f[2][:, 1:].any()
    alone = eng.energy_forces(_batch([base], [types[3]]))
    assert alone["energy"].cpu().numpy()[0] == e[3] and np.array_equal(alone["forces"].cpu().numpy(), f[3])
    small = engine.ClassicalEngine(eng.kind, eng.params.cpu().numpy(), eng.ntypes, n_max=len(base["numbers"]) - 1, max_nbr=32)
    with pytest.raises(_lib.VssrError):
        small.energy_forces(_batch([base], [types[3]]))
