"""The oracle is only trusted because it reproduces the reference's own printed numbers."""
import numpy as np
import pytest
import torch

from oracle import classical, relax
from oracle.nbrlist import neighbor_list
from oracle.painn import EnsembleOracle, PainnOracle, init_random_weights, surface_energy
from oracle.painn_manual import energy_and_grad_manual

PBC3 = [True, True, True]


@pytest.mark.parametrize("name", ["SrTiO3_001_2x2", "O44Sr12Ti16", "O36Sr12Ti12", "O40Sr16Ti12"])
def test_painn_ensemble_golden(structures, potentials, golden_values, sto_weights, name):
    g = golden_values["painn_ensemble_step0"][name]
    s = structures[name]
    ens = EnsembleOracle(sto_weights, potentials["offset_data"], dtype=torch.float64)
    r = ens.calculate(s["positions"], s["numbers"], s["cell"], PBC3)
    # CIF-derived structures carry 5-decimal coordinates: ~1e-4 eV noise
    tol = 2e-5 if name == "SrTiO3_001_2x2" else 2e-4
    assert abs(r["energy"][0] - g["energy"]) < tol
    fn = np.linalg.norm(r["forces"], axis=1)
    if "free" in g:
        assert abs(fn[g["free"]].max() - g["fmax"]) < 2e-5
    else:
        assert np.abs(fn - g["fmax"]).min() < 1e-5   # some atom carries exactly the logged fmax


def test_neighbor_counts(structures, golden_values):
    s = structures["SrTiO3_001_2x2"]
    i, j, S = neighbor_list(s["positions"], s["cell"], PBC3, 6.0)
    assert len(i) == golden_values["edge_counts_sto_2x2"]["cutoff6"]
    i5, _, _ = neighbor_list(s["positions"], s["cell"], PBC3, 5.0)
    assert len(i5) == golden_values["edge_counts_sto_2x2"]["cutoff5"]
    # directed and symmetric: every (i,j,S) has (j,i,-S)
    fwd = set(zip(i.tolist(), j.tolist(), map(tuple, S.tolist())))
    assert all((b, a, (-s0, -s1, -s2)) in fwd for a, b, (s0, s1, s2) in fwd)


def test_bfgs_log_and_surface_energy(structures, potentials, golden_values, sto_weights):
    s = structures["SrTiO3_001_2x2"]
    gold = golden_values["bfgs_log_pristine_sto"]
    ens = EnsembleOracle(sto_weights, potentials["offset_data"], dtype=torch.float32)
    fixed = relax.fixed_mask_from_surface_depth(s["positions"], s["cell"], 1)
    assert np.where(~fixed)[0].tolist() == golden_values["painn_ensemble_step0"]["SrTiO3_001_2x2"]["free"]
    nb = ens.build_nbrs(s["positions"], s["cell"], PBC3)

    def calc(x):
        r = ens.calculate(x, s["numbers"], s["cell"], PBC3, nb)
        return r["energy"][0], r["forces"]

    log = []
    out = relax.relax(calc, s["positions"], fixed, optimizer="BFGS", relax_steps=20, log=log)
    assert out["converged"] and out["nsteps"] == 4
    assert np.allclose([l[1] for l in log], gold["energy"], atol=1e-4)
    assert np.allclose([l[2] for l in log], gold["fmax"], atol=2e-5)
    se = surface_energy(out["energy"], s["numbers"], potentials["offset_data"],
                        golden_values["pristine_sto_surface_energy"]["chem_pots"])
    assert abs(se - golden_values["pristine_sto_surface_energy"]["value"]) < 1e-3


def test_tersoff_golden(structures, potentials, golden_values):
    s = structures["GaN_0001_3x3"]
    prm = classical.TersoffParams(potentials["GaN.tersoff"], ["Ga", "N"])
    types = torch.tensor([0 if z == 31 else 1 for z in s["numbers"]])
    e, f = classical.energy_forces(classical.tersoff_energy, s["positions"], types, s["cell"], PBC3, prm)
    assert abs(e - golden_values["tersoff_gan_pristine"]["energy"]) < 1e-3


def test_manual_backward_matches_autograd(structures, sto_weights):
    s = structures["O40Sr16Ti12"]
    rng = np.random.default_rng(0)
    pos = s["positions"] + rng.normal(0, 0.05, s["positions"].shape)
    i, j, S = neighbor_list(pos, s["cell"], PBC3, 6.0)
    off = S.astype(np.float64) @ s["cell"]
    for st in (sto_weights[0], init_random_weights(1)):
        e, g = PainnOracle(st, dtype=torch.float64).energy_and_grad(pos, s["numbers"], i, j, off)
        e2, g2 = energy_and_grad_manual(st, pos, s["numbers"], i, j, off)
        assert abs(float(e) - float(e2)) < 1e-9
        assert float((g - g2).abs().max()) < 1e-8


def test_eam_au_golden(structures, golden_values):
    """tests/test_Au.py:19 of the reference: min over the 28 canonical states (6 of 8 adatoms)."""
    import itertools
    from pathlib import Path
    from oracle.eam import EAMFuncfl
    z = np.load(Path(__file__).resolve().parent / "golden" / "eam_funcfl.npz")
    au = EAMFuncfl({k.split("/")[1]: z[k] for k in z.files if k.startswith("Au/")})
    slab = structures["Au_110_2x2"]
    ads = structures["Au_110_2x2_proper_adsorbed"]["positions"][16:24]
    best = min(au.energy_forces(np.concatenate([slab["positions"], ads[list(keep)]]), slab["cell"], slab["pbc"])[0]
               for keep in itertools.combinations(range(8), 6))
    assert abs(best - golden_values["eam_au"]["energy"]) < 1e-5     # CIF adatoms carry 5 decimals
    cu = EAMFuncfl({k.split("/")[1]: z[k] for k in z.files if k.startswith("Cu/")})
    a = 3.615
    e, f = cu.energy_forces(np.array([[0, 0, 0], [0, .5, .5], [.5, 0, .5], [.5, .5, 0]]) * a, np.eye(3) * a, [True] * 3)
    assert abs(e / 4 + 3.54) < 1e-6 and np.abs(f).max() < 1e-10   # Foiles-86 Cu: E_coh = 3.54 eV by construction


def test_pourbaix_formula_product_vs_oracle():
    """H7: the host-side scalar of NFFPourbaix against the oracle restatement (no GPU needed: the slab
    energy is injected)."""
    from oracle.pourbaix import pourbaix_potential
    from surface_sampling_b200.atoms import Atoms
    from surface_sampling_b200.calculators import NFFPourbaix, PourbaixAtom
    table = {   # Sr / O literals: reference tests/pourbaix/test_pourbaix_atoms.py:44-86; Ti, H synthetic
        "Sr": dict(E_std=-1.6895, dG2_std=-5.798, n_e=2, n_H=0, conc=1e-6),
        "O": dict(E_std=-4.9480, dG2_std=2.46, n_e=-2, n_H=-2, conc=1.0),
        "Ti": dict(E_std=-7.8955, dG2_std=-9.20, n_e=4, n_H=4, conc=1e-6),
        "H": dict(E_std=-3.3927, dG2_std=0.0, n_e=1, n_H=1, conc=1.0),
    }
    calc = NFFPourbaix.__new__(NFFPourbaix)   # host-only use: no engine, no device
    calc.parameters, calc.results, calc.atoms, calc._cache_key = {}, {}, None, None
    calc.pourbaix_atoms = {k: PourbaixAtom(k, species_conc=v["conc"], num_e=v["n_e"], num_H=v["n_H"],
                                           atom_std_state_energy=v["E_std"], delta_G2_std=v["dG2_std"])
                           for k, v in table.items()}
    for symbols, corr in ((["Sr"] * 4 + ["Ti"] * 4 + ["O"] * 13, {}),
                          (["Sr"] * 2 + ["Ti"] * 2 + ["O"] * 8 + ["H"] * 3, {"HO": 0.23}),
                          (["Sr"] * 2 + ["Ti"] * 2 + ["O"] * 7 + ["H"] * 9, {"HO": 0.23})):   # excess-H water rule
        atoms = Atoms(symbols=symbols, positions=np.zeros((len(symbols), 3)))
        for phi, pH in ((0.0, 0.0), (1.0, 7.0), (-0.5, 14.0)):
            calc.phi, calc.pH, calc.temp, calc.adsorbate_corrections = phi, pH, 0.0257, corr
            e_slab = -123.456
            got = -(calc.get_delta_G1(atoms, slab_energy=e_slab) + calc.get_delta_G2(atoms))
            ref = pourbaix_potential(symbols, e_slab, table, phi, pH, 0.0257, corr)
            assert abs(got - ref) < 1e-10, (symbols, phi, pH, got, ref)
            # the batched driver's scalar (one chain per (pH, phi) grid point) is the same function
            assert calc.surface_energy_fn(phi, pH)(e_slab, symbols) == pytest.approx(ref, abs=1e-10)
            assert calc.surface_energy_fn()(e_slab, symbols) == pytest.approx(got, abs=1e-12)


@pytest.mark.parametrize("name", ["O44Sr12Ti16", "O36Sr12Ti12", "O40Sr16Ti12"])
def test_bfgs_logs_of_the_reference_slabs(structures, potentials, golden_values, sto_weights, name):
    """tests/test_SrTiO3_terms.ipynb:201-230,272-274: three more BFGS logs (6, 3 and 14 lines) and the relaxed surface
    energies 35.931 / 12.478 / -4.876 eV.  Pins ensemble + FixAtoms + CatKit layer tags + BFGS + irun + mu-offset, both
    for the oracle's BFGS and for the PRODUCT's host BFGS (dynamics.HostBFGS, free-atom subspace) driven by the
    oracle's forces."""
    from surface_sampling_b200 import dynamics, engine
    gold = golden_values["bfgs_logs_ref_slabs"][name]
    s = structures[name]
    ens = EnsembleOracle(sto_weights, potentials["offset_data"], dtype=torch.float32)
    fixed = relax.fixed_mask_from_surface_depth(s["positions"], s["cell"], 1)
    assert np.where(~fixed)[0].tolist() == gold["free"]
    nb = ens.build_nbrs(s["positions"], s["cell"], PBC3)

    def calc(x):
        r = ens.calculate(x, s["numbers"], s["cell"], PBC3, nb)
        return r["energy"][0], r["forces"]

    log = []
    out = relax.relax(calc, s["positions"], fixed, optimizer="BFGS", relax_steps=20, log=log)
    assert out["converged"] and len(log) == len(gold["energy"])
    # CIF coordinates carry 5 decimals: ~1e-4 eV, ~1e-4 eV/A noise against the printed log
    assert np.allclose([l[1] for l in log], gold["energy"], atol=2.5e-4)
    assert np.allclose([l[2] for l in log], gold["fmax"], atol=1e-4)
    chem = golden_values["pristine_sto_surface_energy"]["chem_pots"]
    assert abs(surface_energy(out["energy"], s["numbers"], potentials["offset_data"], chem) - gold["surface_energy"]) < 1e-3

    class OracleEngine:      # stands in for PainnEngine: same call, oracle physics (this is a CPU test)
        cutoff, skin = 5.0, 1.0

        def energy_forces(self, batch, z_host=None, nbrs=None):
            r = ens.calculate(batch.pos.numpy(), s["numbers"], s["cell"], PBC3, nb)
            return {"energy": torch.tensor(r["energy"]), "energy_std": torch.tensor(r["energy_std"]),
                    "forces": torch.tensor(r["forces"])}

    b = engine.Batch.from_arrays([s["positions"]], [s["numbers"]], [s["cell"]], [PBC3], [fixed], device="cpu", pinned=False)
    plog = []
    orig = dynamics.eng.neighbor_list
    dynamics.eng.neighbor_list = lambda *a, **k: None
    try:
        o2, f2 = dynamics.relax_host_batch(OracleEngine(), b, s["numbers"], "BFGS", 20, 0.01,
                                           observer=lambda st, e, f, p: plog.append((st, float(e))))
    finally:
        dynamics.eng.neighbor_list = orig
    assert [q[0] for q in plog] == list(range(len(gold["energy"])))
    assert np.allclose([q[1] for q in plog], [l[1] for l in log], atol=2e-5)       # fp32 energies, different eigh sizes
    assert o2[0, 4] == out["nsteps"] and o2[0, 5] == 1.0 and abs(o2[0, 2] - out["raw_energy"]) < 2e-5
    assert np.abs(b.pos.numpy() - out["pos"]).max() < 1e-5


def cu100_slab(a=3.6147, vacuum=15.0):
    """catkit.build.surface(bulk('Cu','fcc',a), size=(2,2,2), miller=(1,0,0), termination=0, vacuum=15) rebuilt
    geometrically (tests/test_Cu.py:29-38; site coordinates logged in tutorials/example.ipynb cell 7: top layer at
    z = 16.807, ontop sites (0,0) (2.556,0) (0,2.556) (2.556,2.556) at z = 18.307)."""
    d, h = a / np.sqrt(2), a / 2
    bottom = [[(i + 0.5) * d, (j + 0.5) * d, vacuum] for i in range(2) for j in range(2)]
    top = [[i * d, j * d, vacuum + h] for i in range(2) for j in range(2)]
    return np.array(bottom + top), np.diag([2 * d, 2 * d, 2 * vacuum + h]), np.array([True, True, False])


def test_eam_cu_golden(golden_values):
    """tests/test_Cu.py:19: min(energy_hist) = -25.2893 is the slab with one Cu on a bridge site."""
    from pathlib import Path
    from oracle.eam import EAMFuncfl
    z = np.load(Path(__file__).resolve().parent / "golden" / "eam_funcfl.npz")
    cu = EAMFuncfl({k.split("/")[1]: z[k] for k in z.files if k.startswith("Cu/")})
    g = golden_values["eam_cu"]
    pos, cell, pbc = cu100_slab(g["a"])
    d = g["a"] / np.sqrt(2)
    bridge = np.array([0.5 * d, 0.0, pos[:, 2].max() + g["planar_distance"]])
    e, _ = cu.energy_forces(np.vstack([pos, bridge]), cell, pbc)
    assert np.allclose(e, g["energy"])          # the reference's own assertion
    assert abs(e - g["energy"]) < 1e-4
    ontop = np.array([0.0, 0.0, pos[:, 2].max() + g["planar_distance"]])
    assert cu.energy_forces(np.vstack([pos, ontop]), cell, pbc)[0] > e      # ontop at 1.5 A is repulsive
