#!/usr/bin/env python
"""Generate tests/golden/mc_painn_oracle_chains.json: single-chain ORACLE VSSR-MC runs on the headline workload.

Each chain is oracle.mc.run_chain (the reference's MC loop with its global-RNG call order) whose energy function is
oracle FIRE (relax_steps=20, fmax=0.01, surface_depth=1 FixAtoms) over the fp32 oracle PaiNN ensemble with the
reference's REAL checkpoints (tests/golden/painn_sto_weights.npz) + the mu-corrected surface energy.  The -m gpu test
tests/test_gpu_mc_painn.py replays the same seeds through MultiChainMC + PainnEngine.relax (memo + constrained
gradients, as benched) and demands identical accept flags, uniforms and occupancies.

A decision whose Boltzmann margin lies inside the energy tolerance cannot be required to agree between two fp32
implementations (north star: 1e-5 eV/atom); the generator records the margin |dE + T ln u| of every decision and only
keeps seeds whose smallest margin exceeds BAND_FACTOR x 1e-5 eV/atom x n_atoms (rejected seeds are listed too).

Runs on CPU only (no reference tree needed):  python tests/golden/make_mc_fixtures.py [n_seeds] [n_procs]
"""
from __future__ import annotations

import json
import multiprocessing as mp
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
GOLD = ROOT / "tests" / "golden"

CONFIG = {
    "slab": "SrTiO3_001_2x2", "n_sites": 64, "site_height": 1.5, "adsorbates": ["Sr", "Ti", "O"],
    "chem_pots": {"Sr": -2, "Ti": 0, "O": 0}, "surface_depth": 1, "relax_steps": 20, "fmax": 0.01,
    "total_sweeps": 4, "sweep_size": 10, "start_temp": 1.0, "alpha": 0.99,
}
E_TOL_PER_ATOM = 1e-5
BAND_FACTOR = 4.0       # curr and prev may each be off by the tolerance, twice over for safety


def run_seed(seed: int) -> dict:
    import torch
    torch.set_num_threads(1)
    from oracle import mc as omc
    from oracle import relax as orelax
    from oracle.painn import EnsembleOracle, load_golden_weights, surface_energy
    from surface_sampling_b200 import mc
    from surface_sampling_b200.engine import NUMBERS, SYMBOLS

    c = CONFIG
    z = np.load(GOLD / "structures.npz")
    pots = json.loads((GOLD / "potentials.json").read_text())
    pos0, num0, cell = z[c["slab"] + "/positions"], z[c["slab"] + "/numbers"], z[c["slab"] + "/cell"]
    pbc = [True, True, True]
    fixed0 = orelax.fixed_mask_from_surface_depth(pos0, cell, c["surface_depth"])
    sites = mc.make_site_grid(pos0, cell, c["n_sites"], c["site_height"])
    ens = EnsembleOracle(load_golden_weights(GOLD / "painn_sto_weights.npz"), pots["offset_data"], dtype=torch.float32)
    n_atoms_log, nsteps_log = [], []

    def energy_fn(symbols, pos):
        num = np.array([NUMBERS[s] for s in symbols])
        fx = np.concatenate([fixed0, np.zeros(len(num) - len(fixed0), bool)])
        nb = ens.build_nbrs(pos, cell, pbc)

        def calc(x):
            r = ens.calculate(x, num, cell, pbc, nb)
            return r["energy"][0], r["forces"]

        o = orelax.relax(calc, pos, fx, optimizer="FIRE", relax_steps=c["relax_steps"], fmax=c["fmax"])
        n_atoms_log.append(len(num))
        nsteps_log.append(o["nsteps"])
        # the OOB clamp is invisible to Metropolis (mcmc/system.py:466-469): raw energy
        return surface_energy(o["raw_energy"], num, pots["offset_data"], c["chem_pots"])

    t0 = time.time()
    o = omc.run_chain(seed, [SYMBOLS[int(q)] for q in num0], pos0, sites, c["adsorbates"], energy_fn, c["total_sweeps"],
                      c["sweep_size"], start_temp=c["start_temp"], alpha=c["alpha"])
    temps = [c["start_temp"] * c["alpha"] ** (k // c["sweep_size"]) for k in range(len(o["decisions"]))]
    t, temps = c["start_temp"], []
    for _ in range(c["total_sweeps"]):
        temps += [t] * c["sweep_size"]
        t *= c["alpha"]
    # energy_fn call order: [initial state], then one call per decision
    n_at = n_atoms_log[1:]
    margins = []
    for (acc, curr, prev, u), T in zip(o["decisions"], temps):
        with np.errstate(divide="ignore"):
            margins.append(abs((curr - prev) + T * np.log(u)) if u > 0 else float("inf"))
    return {"seed": seed, "accept": [bool(d[0]) for d in o["decisions"]], "curr": [float(d[1]) for d in o["decisions"]],
            "prev": [float(d[2]) for d in o["decisions"]], "u": [float(d[3]) for d in o["decisions"]],
            "temps": temps, "occ_history": o["occ_history"], "final_occ": o["final"].occ,
            "final_symbols": [a["sym"] for a in o["final"].atoms], "n_atoms": n_at, "fire_steps": nsteps_log[1:],
            "margin_eV": margins, "energy_hist": [float(x) for x in o["energy_hist"]],
            "frac_accept_hist": o["frac_accept_hist"], "ads_hist": o["ads_hist"], "seconds": time.time() - t0}


def main():
    n_seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    n_procs = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    with mp.get_context("spawn").Pool(n_procs) as pool:
        chains = pool.map(run_seed, range(n_seeds), chunksize=1)
    keep, dropped = [], []
    for ch in chains:
        band = [BAND_FACTOR * E_TOL_PER_ATOM * n for n in ch["n_atoms"]]
        bad = [k for k, (m, b) in enumerate(zip(ch["margin_eV"], band)) if m <= b]
        (dropped if bad else keep).append(ch if not bad else {"seed": ch["seed"], "decisions_inside_band": bad,
                                                             "margin_eV": [ch["margin_eV"][k] for k in bad]})
    out = {"_doc": "oracle single-chain VSSR-MC decisions on SrTiO3(001) 2x2 with the real PaiNN checkpoints; generated by "
                   "tests/golden/make_mc_fixtures.py (CPU, fp32 oracle ensemble + oracle FIRE)",
           "config": CONFIG, "e_tol_per_atom": E_TOL_PER_ATOM, "band_factor": BAND_FACTOR, "chains": keep,
           "dropped_seeds": dropped}
    (GOLD / "mc_painn_oracle_chains.json").write_text(json.dumps(out))
    print("kept", [c["seed"] for c in keep], "dropped", [d["seed"] for d in dropped],
          "accept rate %.2f" % np.mean([np.mean(c["accept"]) for c in keep]),
          "max ads", max(max(c["ads_hist"]) for c in keep), "cpu s/chain %.0f" % np.mean([c["seconds"] for c in keep]))


if __name__ == "__main__":
    main()
