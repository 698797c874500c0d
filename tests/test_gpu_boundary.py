"""The drop-in boundary on the GPU: reference-named calculators, optimize_slab, and accept/reject
parity of the multi-chain driver against the single-chain oracle loop with the oracle's physics."""
import copy

import numpy as np
import pytest
import torch

from oracle import classical as ocl
from oracle import mc as omc
from oracle import relax as orelax
from oracle.painn import EnsembleOracle, surface_energy

pytestmark = pytest.mark.gpu
CHEM = {"Sr": -2, "Ti": 0, "O": 0}


def _sto_atoms(structures):
    from surface_sampling_b200.atoms import Atoms, FixAtoms
    s = structures["SrTiO3_001_2x2"]
    fixed = orelax.fixed_mask_from_surface_depth(s["positions"], s["cell"], 1)
    return Atoms(numbers=s["numbers"], positions=s["positions"], cell=s["cell"], pbc=True,
                 constraint=FixAtoms(mask=fixed)), s, fixed


def test_ensemble_nff_surface_calculator(structures, potentials, sto_weights, golden_values):
    from surface_sampling_b200.calculators import EnsembleNFFSurface, get_std_devs_single, get_embeddings_single
    od = potentials["offset_data"]
    calc = EnsembleNFFSurface(sto_weights, offset_data=od)
    changed = calc.set(chem_pots=CHEM, offset_data=od, relax_atoms=True, relax_steps=20, optimizer="FIRE")
    assert set(changed) >= {"chem_pots", "offset_data"} and calc.parameters.copy()["relax_steps"] == 20
    assert "surface_energy" in calc.implemented_properties and len(calc.models) == 3
    atoms, s, _ = _sto_atoms(structures)
    atoms.calc = calc
    e = atoms.get_potential_energy()
    assert e.shape == (1,) and abs(float(e[0]) - golden_values["painn_ensemble_step0"]["SrTiO3_001_2x2"]["energy"]) < 1e-4
    f = atoms.get_forces()                      # FixAtoms applied like ASE
    assert np.abs(np.linalg.norm(f, axis=1).max() - 0.204407) < 2e-5
    se = calc.get_property("surface_energy", atoms=atoms)
    assert abs(float(se) - surface_energy(float(e[0]), s["numbers"], od, CHEM)) < 2e-5   # results are fp32 like NFF
    assert calc.results["forces_std"].shape == (60, 3) and get_std_devs_single(atoms, calc) > 0
    assert get_embeddings_single(atoms, calc).shape == (128,)
    c2 = copy.deepcopy(calc)                    # SurfaceSystem.copy(copy_calc=True)
    assert c2.engine is calc.engine and c2.parameters == calc.parameters
    with pytest.raises(Exception):
        calc.get_property("stress_free_energy", atoms=atoms)


def test_optimize_slab_fused_and_trajectory_agree(structures, potentials, sto_weights):
    from surface_sampling_b200.calculators import EnsembleNFFSurface
    from surface_sampling_b200.dynamics import optimize_slab
    od = potentials["offset_data"]
    calc = EnsembleNFFSurface(sto_weights, offset_data=od)
    calc.set(chem_pots=CHEM, offset_data=od)
    atoms, s, fixed = _sto_atoms(structures)
    atoms.calc = calc
    slab1, traj1, e1, oob1 = optimize_slab(atoms, optimizer="FIRE", save_traj=False, relax_steps=10)
    slab2, traj2, e2, oob2 = optimize_slab(atoms, optimizer="FIRE", save_traj=True, relax_steps=10, record_interval=5)
    assert traj1 is None and not oob1 and not oob2
    assert e1 == e2 and np.array_equal(slab1.get_positions(), slab2.get_positions())   # same kernels, same bits
    assert len(traj2["atoms"]) == 3 and len(traj2["energies"]) == 3                    # steps 0, 5, 10
    assert traj2["energies"][0] > traj2["energies"][-1] and np.all(traj2["forces"][0][fixed] == 0)
    assert np.array_equal(atoms.get_positions(), s["positions"])                       # works on a copy
    # calculator is primed with the final evaluation: no recomputation needed for the surface energy
    se = calc.get_property("surface_energy", atoms=slab1)
    assert abs(se - surface_energy(e1, s["numbers"], od, CHEM)) < 1e-4
    with pytest.raises(NotImplementedError):
        optimize_slab(atoms, optimizer="BFGSLineSearch")


@pytest.mark.parametrize("name", ["SrTiO3_001_2x2", "O44Sr12Ti16", "O36Sr12Ti12", "O40Sr16Ti12"])
def test_optimize_slab_bfgs_reproduces_the_reference_logs(structures, potentials, sto_weights, golden_values, name):
    """The ONLY relax logs the reference tree holds are BFGS (tutorials/SrTiO3_001.ipynb:241-245,
    tests/test_SrTiO3_terms.ipynb:201-230): optimize_slab(optimizer="BFGS") driven by the CUDA ensemble forces reproduces
    each of them line by line (step count, energies, fmax), and the relaxed surface energies 12.471 / 35.931 / 12.478 /
    -4.876 eV (tutorials/SrTiO3_001.ipynb:282, tests/test_SrTiO3_terms.ipynb:272-274)."""
    from surface_sampling_b200.atoms import Atoms, FixAtoms
    from surface_sampling_b200.calculators import EnsembleNFFSurface
    from surface_sampling_b200.dynamics import optimize_slab
    if name == "SrTiO3_001_2x2":
        gold = dict(golden_values["bfgs_log_pristine_sto"], surface_energy=golden_values["pristine_sto_surface_energy"]["value"])
        e_tol, f_tol = 1e-4, 2e-5
    else:
        gold = golden_values["bfgs_logs_ref_slabs"][name]
        e_tol, f_tol = 2.5e-4, 1e-4          # CIF coordinates carry 5 decimals
    s = structures[name]
    od = potentials["offset_data"]
    fixed = orelax.fixed_mask_from_surface_depth(s["positions"], s["cell"], 1)
    calc = EnsembleNFFSurface(sto_weights, offset_data=od)
    calc.set(chem_pots=CHEM, offset_data=od, relax_atoms=True, relax_steps=20, optimizer="BFGS")
    atoms = Atoms(numbers=s["numbers"], positions=s["positions"], cell=s["cell"], pbc=True, constraint=FixAtoms(mask=fixed))
    atoms.calc = calc
    slab, traj, energy, oob = optimize_slab(atoms, optimizer="BFGS", save_traj=True, relax_steps=20, record_interval=1)
    assert not oob and len(traj["energies"]) == len(gold["energy"])          # same number of BFGS lines
    assert np.allclose(traj["energies"], gold["energy"], atol=e_tol), np.abs(np.array(traj["energies"]) - gold["energy"]).max()
    fmax = [np.linalg.norm(f, axis=1).max() for f in traj["forces"]]         # FixAtoms-masked, like the ASE log
    assert np.allclose(fmax, gold["fmax"], atol=f_tol), np.abs(np.array(fmax) - gold["fmax"]).max()
    assert fmax[-1] < 0.01 and abs(energy - gold["energy"][-1]) < e_tol
    assert np.array_equal(slab.get_positions()[fixed], s["positions"][fixed])
    se = calc.get_property("surface_energy", atoms=slab)
    assert abs(float(se) - gold["surface_energy"]) < 1e-3
    # the fused FIRE path relaxes the same slab towards the same minimum (optimizer-dependent: loose check)
    _, _, e_fire, _ = optimize_slab(atoms, optimizer="FIRE", save_traj=False, relax_steps=20)
    assert abs(e_fire - energy) < 0.05


def test_optimize_slab_cg_runs_on_gpu_forces(structures, potentials, sto_weights):
    """optimizer="CG" (ASE SciPyFminCG -> scipy fmin_cg on E/70): API completeness of mcmc/dynamics.py:120-127; unpinned in
    the reference tree, so only the contract is checked: energy decreases, FixAtoms hold, step budget respected."""
    from surface_sampling_b200.atoms import Atoms, FixAtoms
    from surface_sampling_b200.calculators import EnsembleNFFSurface
    from surface_sampling_b200.dynamics import optimize_slab
    s = structures["SrTiO3_001_2x2"]
    fixed = orelax.fixed_mask_from_surface_depth(s["positions"], s["cell"], 1)
    calc = EnsembleNFFSurface(sto_weights, offset_data=potentials["offset_data"])
    atoms = Atoms(numbers=s["numbers"], positions=s["positions"], cell=s["cell"], pbc=True, constraint=FixAtoms(mask=fixed))
    atoms.calc = calc
    e0 = float(atoms.get_potential_energy()[0])
    slab, traj, e, oob = optimize_slab(atoms, optimizer="CG", save_traj=True, relax_steps=10, record_interval=1)
    assert not oob and e < e0 and abs(e - (-467.5413)) < 5e-3
    assert np.array_equal(slab.get_positions()[fixed], s["positions"][fixed]) and 2 <= len(traj["energies"]) <= 11


def test_lammps_surf_calc_and_oob(structures, potentials, golden_values):
    from surface_sampling_b200 import engine
    from surface_sampling_b200.atoms import Atoms
    from surface_sampling_b200.calculators import LAMMPSSurfCalc
    from surface_sampling_b200.dynamics import optimize_slab
    s = structures["GaN_0001_3x3"]
    calc = LAMMPSSurfCalc("tersoff", engine.tersoff_param_table(potentials["GaN.tersoff"], ["Ga", "N"]), ["Ga", "N"],
                          bulk_index=36, n_max=64, max_nbr=24)
    calc.set(relax_steps=50, run_dir=".")
    atoms = Atoms(numbers=s["numbers"], positions=s["positions"], cell=s["cell"], pbc=s["pbc"], calculator=calc)
    e = calc.get_property("surface_energy", atoms=atoms)
    assert abs(e - golden_values["tersoff_gan_pristine"]["energy"]) < 1e-3
    assert abs(calc.results["per_atom_energies"].sum() - e) < 1e-9
    ads = atoms.copy(); ads.append("Ga", s["positions"][35] + np.array([0.3, 0.2, 1.9])); ads.calc = calc
    slab, traj, e_rel, oob = optimize_slab(ads, optimizer="LAMMPS", relax_steps=50)
    assert traj is None and not oob and e_rel < calc.get_potential_energy(atoms=ads)
    assert np.array_equal(slab.get_positions()[:36], s["positions"])       # bulk group frozen
    # overlapping atoms -> |E| or |F| > 1000 -> clamp (dynamics.py:159-168)
    n_idx = int(np.where(s["numbers"] == 7)[0][-1])   # N on top of N: |F| ~ lambda1*A = 3.7e3 eV/A
    bad = atoms.copy(); bad.append("N", s["positions"][n_idx] + np.array([0.05, 0.0, 0.05])); bad.calc = calc
    _, _, e_bad, oob_bad = optimize_slab(bad, optimizer="LAMMPS", relax_steps=0)
    assert oob_bad and e_bad == 1000


def test_accept_reject_parity_tersoff(structures, potentials):
    """Same seeds -> the multi-chain GPU run makes exactly the decisions of the single-chain oracle loop
    (oracle Tersoff + oracle FIRE, fp64)."""
    from surface_sampling_b200 import engine, mc
    s = structures["GaN_0001_3x3"]
    tab = engine.tersoff_param_table(potentials["GaN.tersoff"], ["Ga", "N"])
    eng = engine.ClassicalEngine(engine.POT_TERSOFF, tab, 2, n_max=64, max_nbr=24)
    prm = ocl.TersoffParams(potentials["GaN.tersoff"], ["Ga", "N"])
    sites = mc.make_site_grid(s["positions"], s["cell"], 12, 1.8)
    fixed0 = np.ones(36, bool)
    tmap = {31: 0, 7: 1}
    steps = 15

    def relax_fn(pos_l, num_l, fix_l):
        types = [np.array([tmap[int(z)] for z in zz], np.int32) for zz in num_l]
        b = engine.Batch.from_arrays(pos_l, types, [s["cell"]] * len(pos_l), [s["pbc"]] * len(pos_l), fix_l)
        return eng.relax(b, relax_steps=steps)["out"].cpu().numpy()

    seeds = [0, 5]
    drv = mc.MultiChainMC(s["numbers"], s["positions"], fixed0, sites, ["Ga"], relax_fn, lambda e, sym: e, seeds)
    res = drv.run(total_sweeps=2, sweep_size=4, start_temp=0.5, perform_annealing=True, alpha=0.9)

    def energy_fn(symbols, pos):
        types = torch.tensor([0 if q == "Ga" else 1 for q in symbols])
        fx = np.arange(len(symbols)) < 36
        o = orelax.relax(lambda x: ocl.energy_forces(ocl.tersoff_energy, x, types, s["cell"], s["pbc"], prm), pos, fx,
                         optimizer="FIRE", relax_steps=steps)
        return o["raw_energy"]

    sym0 = ["Ga" if z == 31 else "N" for z in s["numbers"]]
    for k, sd in enumerate(seeds):
        o = omc.run_chain(sd, sym0, s["positions"], sites, ["Ga"], energy_fn, 2, 4, start_temp=0.5, alpha=0.9)
        assert [d[0] for d in drv.decisions[k]] == [d[0] for d in o["decisions"]]
        assert np.allclose([d[1] for d in drv.decisions[k]], [d[1] for d in o["decisions"]], rtol=1e-9, atol=1e-7)
        assert list(drv.chains[k].occ) == o["final"].occ
        # no decision sat inside the energy-tolerance band of its uniform draw
        assert all(abs(np.exp(-(c - p) / 0.5) - u) > 1e-6 or c <= p for _, c, p, u in o["decisions"])


POURBAIX_TABLE = {   # Sr / O / H: literals of the reference's tests/pourbaix/test_pourbaix_atoms.py:44-86; Ti synthetic
    "Sr": dict(E_std=-1.68949, dG2_std=-5.79807, n_e=2, n_H=0, conc=1e-6),
    "Ti": dict(E_std=-7.8955, dG2_std=-9.20, n_e=4, n_H=4, conc=1.0),
    "O": dict(E_std=-5.26469, dG2_std=-2.45830, n_e=-2, n_H=-2, conc=1.0),
    "H": dict(E_std=-4.0356, dG2_std=0.0, n_e=1, n_H=1, conc=1.0),
}


def test_nff_pourbaix_calculator_on_gpu(structures, sto_weights):
    """H7 through the product calculator on the GPU: NFFPourbaix(models[0]).calculate -> results['surface_energy'] ==
    results['pourbaix_potential'] == oracle formula on the oracle's single-model energy (scripts/
    sample_pourbaix_surface.py:253-259, calculators.py:197-361), for a slab carrying an HO group (adsorbate correction +
    excess-H rule), at several (phi, pH)."""
    from oracle.pourbaix import pourbaix_potential
    from surface_sampling_b200.atoms import Atoms
    from surface_sampling_b200.calculators import NFFPourbaix, PourbaixAtom
    s = structures["SrTiO3_001_2x2"]
    ztop = s["positions"][:, 2].max()
    ads_pos = np.array([[1.0, 1.0, ztop + 1.6], [2.0, 1.0, ztop + 1.6], [5.0, 4.0, ztop + 1.7]])
    numbers = np.concatenate([s["numbers"], [8, 1, 38]])
    pos = np.vstack([s["positions"], ads_pos])
    atoms = Atoms(numbers=numbers, positions=pos, cell=s["cell"], pbc=True)
    patoms = {k: PourbaixAtom(k, species_conc=v["conc"], num_e=v["n_e"], num_H=v["n_H"], atom_std_state_energy=v["E_std"],
                              delta_G2_std=v["dG2_std"]) for k, v in POURBAIX_TABLE.items()}
    calc = NFFPourbaix(sto_weights[0], device="cuda", model_units="kcal/mol", prediction_units="eV")
    assert len(calc.models) == 1
    ens1 = EnsembleOracle(sto_weights[:1], None, dtype=torch.float64)
    e_ref = ens1.calculate(pos, numbers, s["cell"], [True] * 3)["energy"][0]
    symbols = atoms.get_chemical_symbols()
    for phi, pH in ((1.0, 12.0), (0.0, 0.0), (-0.5, 7.0)):
        changed = calc.set(temperature=0.0257, phi=phi, pH=pH, pourbaix_atoms=patoms, adsorbate_corrections={"HO": 0.23})
        assert "phi" in changed or "pH" in changed or "temperature" in changed
        atoms.calc = calc
        calc.calculate(atoms)
        r = calc.results
        ref = pourbaix_potential(symbols, e_ref, POURBAIX_TABLE, phi, pH, 0.0257, {"HO": 0.23})
        assert abs(float(r["energy"][0]) - e_ref) < 1e-5 * len(numbers)
        assert r["surface_energy"] == r["pourbaix_potential"]
        assert abs(r["surface_energy"] - ref) < 1e-5 * len(numbers), (phi, pH, r["surface_energy"], ref)
        assert abs(calc.get_property("surface_energy", atoms=atoms) - r["surface_energy"]) == 0.0       # cached
        # the batched driver's scalar is the same function of (energy, symbols)
        # (the calculator's results are fp32 like NFF's -- E_slab + correction stays a float32 array -- the driver's fp64)
        assert abs(calc.surface_energy_fn()(float(r["energy"][0]), symbols) - float(r["surface_energy"][0])) < 1e-4


def test_embedding_values_vs_oracle(structures, potentials, sto_weights):
    """`embedding` = final scalar features s^(3) (what get_embeddings_single feeds the latent-space
    clustering), compared VALUE BY VALUE with the oracle's return_features, per model and ensemble mean."""
    from oracle.painn import PainnOracle
    from surface_sampling_b200 import engine
    from surface_sampling_b200.atoms import Atoms
    from surface_sampling_b200.calculators import EnsembleNFFSurface, get_embeddings_single
    s = structures["SrTiO3_001_2x2"]
    eng = engine.PainnEngine(sto_weights, potentials["offset_data"])
    b = engine.Batch.from_arrays([s["positions"], s["positions"] + 0.01], [s["numbers"]] * 2, [s["cell"]] * 2, [[True] * 3] * 2)
    emb = eng.energy_forces(b, want_embedding=True)["embedding"].cpu().numpy()        # [M, A, 128]
    assert emb.shape == (3, 120, 128)
    ens = EnsembleOracle(sto_weights, None, dtype=torch.float64)
    i, j, S, off = ens.build_nbrs(s["positions"], s["cell"], [True] * 3)
    feats = []
    for m, st in enumerate(sto_weights):
        o = PainnOracle(st, dtype=torch.float64)
        _, f, _ = o.forward_energy(torch.tensor(s["positions"]), torch.tensor(s["numbers"]).long(), torch.tensor(i).long(),
                                   torch.tensor(j).long(), torch.tensor(off), return_features=True)
        feats.append(f.numpy())
        scale = np.abs(feats[-1]).max()
        assert np.abs(emb[m, :60] - feats[-1]).max() < 2e-5 * max(scale, 1.0), (m, np.abs(emb[m, :60] - feats[-1]).max(), scale)
    calc = EnsembleNFFSurface(sto_weights, offset_data=potentials["offset_data"])
    atoms = Atoms(numbers=s["numbers"], positions=s["positions"], cell=s["cell"], pbc=True)
    e = get_embeddings_single(atoms, calc)                  # mean over models, then over atoms -> [128]
    ref = np.mean(feats, axis=0).mean(axis=0)
    assert e.shape == (128,) and np.abs(e - ref).max() < 2e-5 * max(np.abs(ref).max(), 1.0)
