#!/usr/bin/env python
"""bench.py — relaxed VSSR-MC proposals/s (headline: SrTiO3(001) 2x2, PaiNN 3-model ensemble, BASELINE configs[3]).

A "step" is one MC iteration of every chain on this GPU: propose (host, reference RNG order) -> ideal-site structure ->
H2D -> neighbour list + FIRE relaxation with ensemble force evaluations (GPU, no host round trip) -> 8 scalars per chain
D2H -> Metropolis accept/reject.  Chains are first BURNT IN (--burn-in MC steps through the same public driver) so that
the timed region sees the adsorbate coverage of a running chain, not a pristine slab; the coverage is printed in `config`.

  value        relaxed proposals/s with the step's inputs already resident in HBM (relax call only)
  e2e          the same metric through the public API (MultiChainMC.pipeline) with HOST buffers: pinned H2D of
               positions/species/masks and D2H of the result inside the timed region
  roofline     dominant kernel class, timed live with CUDA-event pairs on the launching stream; per class both the
               ALGORITHMIC work and the work the kernels EXECUTE (memoised edges skip the filter) against
               the peak of the pipe they run on
  cpu_baseline the CPU oracle port of the reference path on the host cores (bounded sample)
  reference_equivalent_gpu   the same torch restatement run eagerly on this B200, single chain
  workloads    short runs of the other BASELINE configs (GaN Tersoff, Si SW, SrTiO3 Pourbaix grid), each with its own
               value / e2e / roofline / cpu_baseline (skipped with --no-extra-workloads or an explicit --workload)

`--impl reference` times the reference path's CPU restatement (the oracle: the reference's own stack -- ASE/NFF/LAMMPS --
is not installable here) on the same config.  Weak scaling: --chains-per-gpu chains on every rank; chains
(and the (pH, U, chain) units of the Pourbaix grid) never interact, NCCL only gathers per-chain scalars.
"""
from __future__ import annotations

import os
import sys

# torchrun exports OMP_NUM_THREADS=1 to every rank; the reference arm is a CPU measurement "with all the host threads
# it can use", so undo that before numpy / torch initialise their thread pools
if "reference" in sys.argv and os.environ.get("OMP_NUM_THREADS") == "1":
    for _k in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
        os.environ[_k] = str(os.cpu_count() or 1)

import argparse
import contextlib
import json
import subprocess
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
GOLD = ROOT / "tests" / "golden"

CHEM_POTS = {"Sr": -2, "Ti": 0, "O": 0}
FREE = [7, 8, 22, 23, 37, 38, 52, 53]   # surface_depth=1 (tutorials/SrTiO3_001.ipynb cell 7 log)
FFMA2_PEAK = 65.8    # TFLOP/s, packed fp32x2 FMA measured on this pool's B200 (profiles/microbench/ffma2.cu)
FP64_PEAK = 37.0     # TFLOP/s nominal B200 FP64 (148 SMs x 64 DFMA/clk x 2 x 1.965 GHz); not in MEASURED_PEAKS.json
# packed fp32 issue slots (FFMA2 / FMUL2 / FADD2) per edge and lane (= feature pair) of the message kernels, counted in
# csrc/painn_message.cuh: 60 (fwd) / 120 (bwd: w and dw/dd) of them are the radial filter W_d . rbf
MSG_SLOTS = {"direct_fwd": 72, "direct_bwd": 180, "direct_bwd_frozen_receiver": 73, "memo_fwd": 12, "memo_bwd_state": 13,
             "memo_bwd_full": 60}
GEMM_FLOP_PER_ATOM = 2 * 1491072.0   # node MLPs, forward; the backward (no dW) is the same again
FILTER_FLOP_PER_EDGE = 46080.0       # 3 layers x 2 x 20 x 384, forward; same again for dw/dd

# pH x U grid of BASELINE configs[4]: 8 x 7 = 56 points
PH_VALUES = [0.0, 2.0, 4.0, 6.0, 8.0, 10.0, 12.0, 14.0]
U_VALUES = [-1.0, -0.5, 0.0, 0.5, 1.0, 1.5, 2.0]
# PourbaixAtom table: Sr / O / H rows are the literals of the reference's tests/pourbaix/test_pourbaix_atoms.py:44-86
# (phi=1, pH=0 case); the Ti row is synthetic (TiO2: 4 e-, 4 H+), pymatgen being unavailable
# The reference energies (atom_std_state_energy) are RE-CENTRED for the synthetic model: a random-init PaiNN has no
# DFT-scale atomic energies, so with the literal values adding any atom costs +6 .. +28 eV and nothing is ever accepted
# (the chains would stay pristine).  Shifted by +8.84262 / +20.75244 / +5.89452 / +4.94979 eV, adding one Sr / Ti / O / H
# is grand-potential neutral at the grid centre (pH 7, U 0.5 V); the pH and U slopes (num_e, num_H, species_conc,
# delta_G2_std) stay the reference's, so coverage and species vary over the grid as in a real Pourbaix run.
POURBAIX_TABLE = {
    "Sr": dict(species_conc=1e-6, num_e=2, num_H=0, atom_std_state_energy=-1.68949 + 8.84262, delta_G2_std=-5.79807),
    "Ti": dict(species_conc=1.0, num_e=4, num_H=4, atom_std_state_energy=-7.8955 + 20.75244, delta_G2_std=-9.20),
    "O": dict(species_conc=1.0, num_e=-2, num_H=-2, atom_std_state_energy=-5.26469 + 5.89452, delta_G2_std=-2.45830),
    "H": dict(species_conc=1.0, num_e=1, num_H=1, atom_std_state_energy=-4.0356 + 4.94979, delta_G2_std=0.0),
}

# single-chain rates the reference's own notebooks print (other hardware, other optimiser settings)
REFERENCE_PUBLISHED = {
    "sto_painn": {"value": 0.0825, "unit": "proposals/s", "what": "SrTiO3(001) 2x2, PaiNN 3-model ensemble, BFGS 20 steps, 1 chain, RTX 2080 Ti",
                  "source": "tutorials/SrTiO3_001.ipynb:1558"},
    "gan_tersoff": {"value": 31.8, "unit": "proposals/s", "what": "GaN(0001) 3x3, Tersoff, LAMMPS CG <=100 its, 1 chain, CPU",
                    "source": "tutorials/GaN_0001.ipynb:6356"},
}

WORKLOADS = {
    "sto_painn": dict(slab="SrTiO3_001_2x2", n_sites=64, adsorbates=["Sr", "Ti", "O"], relax_steps=20, canonical=False,
                      num_ads=0, height=1.5, chains=128, temp=1.0, burn_in=200, cpu_props=16, models=3,
                      desc="SrTiO3(001) 2x2 VSSR-MC, PaiNN 3-model ensemble (random-init weights seeds 0,1,2), semigrand "
                           "Sr/Ti/O on 64 virtual sites, FIRE relax_steps=20 fmax=0.01 (BASELINE.json configs[3])"),
    "gan_tersoff": dict(slab="GaN_0001_3x3", n_sites=107, adsorbates=["Ga"], relax_steps=100, canonical=True, num_ads=12,
                        height=1.8, chains=256, temp=1.0, burn_in=20, cpu_props=2, models=0,
                        desc="GaN(0001) 3x3 VSSR-MC, Tersoff (Nord 2003), canonical 12 Ga adatoms on 107 virtual sites, "
                             "FIRE relax_steps=100 fmax=0.01, bulk ids<=36 frozen (BASELINE.json configs[1])"),
    "si_sw": dict(slab="Si_111_5x5", n_sites=100, adsorbates=["Si"], relax_steps=100, canonical=False, num_ads=0,
                  height=2.0, chains=256, temp=1.0, burn_in=20, cpu_props=3, models=0,
                  desc="Si(111) 5x5 VSSR-MC, Stillinger-Weber (SW-1985 literature parameters, parity unpinned), semigrand "
                       "Si on 100 virtual sites, FIRE relax_steps=100 fmax=0.01, ids<=75 frozen (BASELINE.json configs[2])"),
    "sto_pourbaix": dict(slab="SrTiO3_001_2x2", n_sites=64, adsorbates=["Sr", "Ti", "O", "HO"], relax_steps=20,
                         canonical=False, num_ads=0, height=1.5, chains=112, temp=0.257, burn_in=150, cpu_props=3, models=1,
                         desc="SrTiO3(001) 2x2 Pourbaix VSSR-MC (sample_pourbaix_surface.py): single PaiNN model (random-init, "
                              "seed 0) + NFFPourbaix grand potential, 8 pH x 7 U grid points x chains sharded as (pH,U,chain) "
                              "units, semigrand Sr/Ti/O/HO on 64 sites, HO correction 0.23 eV, kT=0.0257, T=0.257, reference energies "
                              "re-centred for the random-init model (grand-potential neutral at pH 7, U 0.5 V), FIRE "
                              "relax_steps=20 (BASELINE.json configs[4])"),
}


def load_workload(name):
    w = WORKLOADS[name]
    z = np.load(GOLD / "structures.npz")
    pots = json.loads((GOLD / "potentials.json").read_text())
    n = w["slab"]
    pos, num, cell, pbc = z[f"{n}/positions"], z[f"{n}/numbers"], z[f"{n}/cell"], z[f"{n}/pbc"]
    if name in ("sto_painn", "sto_pourbaix"):
        fixed = np.ones(len(num), bool)
        fixed[FREE] = False
        pbc = np.array([True, True, True])
    elif name == "gan_tersoff":
        fixed = np.ones(len(num), bool)                 # `group bulk id <= 36`
    else:
        fixed = np.arange(len(num)) < 75                # `group bulk id <= 75`
    return pos, num, cell, pbc, fixed, pots


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, index: int, enabled: bool = True):
        self.index, self.proc, self.lines, self.enabled = index, None, [], enabled

    def start(self):
        if not self.enabled:
            return
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["not sampled on this rank" if not self.enabled
                                                                     else "nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            t = [x.strip() for x in l.split(",")]
            if len(t) < 7:
                continue
            try:
                sm.append(float(t[0])); mx.append(float(t[1])); pw.append(float(t[2]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if t[3 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------- host scalars
def pourbaix_calc():
    """Host-only NFFPourbaix (no engine): the table / temperature / corrections of the Pourbaix workload."""
    from surface_sampling_b200.calculators import NFFPourbaix, PourbaixAtom
    calc = NFFPourbaix.__new__(NFFPourbaix)
    calc.parameters, calc.results, calc.atoms, calc._cache_key = {}, {}, None, None
    calc.pourbaix_atoms = {k: PourbaixAtom(k, **v) for k, v in POURBAIX_TABLE.items()}
    calc.temp, calc.phi, calc.pH, calc.adsorbate_corrections = 0.0257, 0.0, 7.0, {"HO": 0.23}
    return calc


def grid_units(name, C, rank, world):
    """(pH, U, chain) units of this rank for the Pourbaix grid (parallel.shard_grid); None for the other workloads."""
    if name != "sto_pourbaix":
        return None
    from surface_sampling_b200.parallel import shard_grid
    n_points = len(PH_VALUES) * len(U_VALUES)
    cpp = max(1, round(C * world / n_points))           # chains per grid point over the whole job
    units = [(ph, u, c) for ph in PH_VALUES for u in U_VALUES for c in range(cpp)]
    mine = shard_grid(PH_VALUES, U_VALUES, cpp, rank, world)
    seeds = [units.index(t) for t in mine]              # global unit id = chain seed
    return mine, seeds, cpp


def surface_energy_fns(name, pots, units):
    if name in ("gan_tersoff", "si_sw"):
        return lambda e, sym: e           # LAMMPSSurfCalc: surface energy = potential energy
    if name == "sto_pourbaix":
        calc = pourbaix_calc()
        return [calc.surface_energy_fn(phi=u, pH=ph) for ph, u, _ in units]
    from surface_sampling_b200.calculators import surface_energy_from
    od = pots["offset_data"]
    return lambda e, sym: surface_energy_from(e, sym, od, CHEM_POTS)


def build_driver(name, relax_fn, seeds, units=None):
    from surface_sampling_b200 import mc
    pos, num, cell, pbc, fixed, pots = load_workload(name)
    w = WORKLOADS[name]
    sites = mc.make_site_grid(pos, cell, w["n_sites"], w["height"])
    drv = mc.MultiChainMC(num, pos, fixed, sites, w["adsorbates"], relax_fn, surface_energy_fns(name, pots, units), seeds,
                          canonical=w["canonical"], num_ads_atoms=w["num_ads"])
    drv.temp = w["temp"]
    return drv


# ---------------------------------------------------------------------------------------------- CPU / torch baselines
def make_oracle_relax_fn(name, evals, device="cpu"):
    """Single-chain restatement of the hot path for `name` (oracle physics + oracle FIRE); device='cuda' runs the same
    eager torch code on the GPU (the "reference-equivalent GPU path")."""
    import torch
    from oracle import classical as ocl
    from oracle import relax as orelax
    from oracle.painn import EnsembleOracle
    from surface_sampling_b200.loaders import init_random_weights

    pos0, num0, cell, pbc, fixed0, pots = load_workload(name)
    w = WORKLOADS[name]
    if w["models"]:
        ens = EnsembleOracle([init_random_weights(s) for s in range(w["models"])],
                             pots["offset_data"] if name == "sto_painn" else None, dtype=torch.float32, device=device)

        def energy_forces_factory(p, zz):
            nb = ens.build_nbrs(p, cell, pbc)

            def calc(x):
                r = ens.calculate(x, zz, cell, pbc, nb)
                evals[0] += w["models"] * len(zz)
                return r["energy"][0], r["forces"]
            return calc
    elif name == "gan_tersoff":
        prm = ocl.TersoffParams(pots["GaN.tersoff"], ["Ga", "N"])

        def energy_forces_factory(p, zz):
            types = torch.tensor([0 if q == 31 else 1 for q in zz])
            return lambda x: ocl.energy_forces(ocl.tersoff_energy, x, types, cell, pbc, prm)
    else:
        def energy_forces_factory(p, zz):
            return lambda x: ocl.energy_forces(ocl.sw_energy, x, cell, pbc, ocl.SWParams())

    def relax_fn(pos_l, num_l, fix_l):
        out = np.zeros((len(pos_l), 8))
        for k, (p, zz, fx) in enumerate(zip(pos_l, num_l, fix_l)):
            o = orelax.relax(energy_forces_factory(p, zz), p, fx, optimizer="FIRE", relax_steps=w["relax_steps"], fmax=0.01)
            out[k, 0], out[k, 2], out[k, 4] = o["energy"], o["raw_energy"], o["nsteps"]
        return out
    return relax_fn


def oracle_proposals(name: str, n_proposals: int, threads: int, seed0: int = 0, device="cpu", burn_occ=None):
    """Reference-path restatement, single chain.  Returns (proposals, seconds, atom_model_evals).  `burn_occ` seeds the
    chain with an adsorbate coverage (list of (site, adsorbate)) so the sample is taken in the benched regime."""
    import torch
    torch.set_num_threads(threads)
    evals = [0]
    units = [(PH_VALUES[3], U_VALUES[2], 0)] if name == "sto_pourbaix" else None
    drv = build_driver(name, make_oracle_relax_fn(name, evals, device), [seed0], units)
    for site, ads in (burn_occ or []):
        drv.chains[0].change_site(site, ads)
    if WORKLOADS[name]["canonical"]:
        drv.prepare_canonical()
    drv._ensure_prev(drv.chains)   # the initial-state energy is not a proposal
    drv.n_relaxed, evals[0] = 0, 0
    if device != "cpu":
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n_proposals):
        drv.step()
    if device != "cpu":
        torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return drv.n_relaxed, dt, evals[0]


def typical_occupancy(drv, k=0):
    """(site, adsorbate) list of the chain whose adsorbate count is the median of the burnt-in population."""
    from surface_sampling_b200.mc import hill_formula
    order = np.argsort([c.num_adsorbates for c in drv.chains])
    c = drv.chains[int(order[len(order) // 2])]
    return [(int(s), hill_formula(c.symbols_at_site(int(s)))) for s in np.flatnonzero(c.occ)]


def workload_config(name, args, chains, world, extra=None):
    cfg = {"workload": WORKLOADS[name]["desc"], "chains_per_gpu": chains,
           "l2": "per-evaluation working set (activations ~48 KB/atom/model, >1 GB) exceeds the 126 MB L2; no explicit flush"
                 if WORKLOADS[name]["models"] else "whole relaxation is shared-memory resident; L2 is not on the path",
           "parallelism": f"chains sharded, {world} rank(s), no data-path collective",
           "e2e_driver": "MultiChainMC.pipeline; PaiNN: 1 chain group per GPU (default stream); classical potentials: "
                         f"{args.groups or 4} interleaved chain groups on separate CUDA streams; value = one batch of all chains per step"}
    if WORKLOADS[name]["models"]:
        cfg["engine_options"] = {
            "filter_memo": not os.environ.get("VSSR_NO_FILTER_MEMO"),
            "constrained_gradients": not (os.environ.get("VSSR_FULL_GRAD") or os.environ.get("VSSR_NO_FILTER_MEMO")),
            "note": "every evaluation runs the full 3-layer forward and backward of all models; the memo holds the radial filter "
                    "rows w(d), dw/dd of frozen-frozen pairs (functions of weights and the frozen geometry only), and dE/dx of "
                    "FixAtoms atoms -- which the optimiser discards -- is not formed during FIRE steps; energies, positions and "
                    "accept/reject are bit-identical to the plain mode (tests/test_gpu_painn.py)"}
    cfg.update(extra or {})
    return cfg


# adsorbates per chain after the default burn-in, measured by the B200 arm on this workload (profiles/round2_bench*.json:
# config.coverage): the reference arm starts its chain at the same coverage so both arms relax structures of the same size
REFERENCE_COVERAGE = {"sto_painn": 32, "sto_pourbaix": 28}


def reference_occupancy(name):
    k = REFERENCE_COVERAGE.get(name, 0)
    w = WORKLOADS[name]
    sites = np.random.RandomState(0).choice(w["n_sites"], size=k, replace=False)
    return [(int(s_), w["adsorbates"][i % len(w["adsorbates"])]) for i, s_ in enumerate(sites)]


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    name = args.workload
    threads = os.cpu_count() or 1
    occ = reference_occupancy(name)
    for _ in range(min(args.warmup, 1)):
        oracle_proposals(name, 1, threads, burn_occ=occ)
    n, dt, evals = 0, 0.0, 0
    for s in range(args.steps):
        a, b, c = oracle_proposals(name, 1, threads, seed0=s, burn_occ=occ)
        n, dt, evals = n + a, dt + b, evals + c
    val = n / dt
    print(json.dumps({
        "impl": "reference", "metric": "relaxed_proposals_per_sec", "value": val, "unit": "proposals/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if WORKLOADS[name]["models"] else "f64", "data": "synthetic",
        "config": workload_config(name, args, 1, 1),
        "cpu_baseline": {"value": val, "unit": "proposals/s", "cores": threads, "kind": "port",
                         "sample": f"{n} single-chain relaxed proposals (1 per step) from a slab carrying {len(occ)} adsorbates -- the "
                                   "coverage the B200 arm reaches after its burn-in -- oracle port of the reference path (torch-CPU "
                                   "physics + numpy FIRE); the reference stack itself is not installable here"},
        "e2e": {"value": val, "unit": "proposals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "painn_atom_model_evals_per_sec": evals / dt if evals else None,
    }))


# ---------------------------------------------------------------------------------------------- the B200 arm
def run_workload(name, args, ctx, steps, warmup, burn_in, headline):
    """One workload on this rank's GPU; returns the JSON line (dict) on every rank (timings are max over ranks)."""
    import torch
    import torch.distributed as dist
    from surface_sampling_b200 import engine, loaders

    lib, world, rank, local = ctx["lib"], ctx["world"], ctx["rank"], ctx["local"]
    w = WORKLOADS[name]
    C = args.chains_per_gpu if (args.chains_per_gpu and headline) else w["chains"]
    steps_relax = w["relax_steps"]
    pos, num, cell, pbc, fixed, pots = load_workload(name)
    grid = grid_units(name, C, rank, world)
    units = None
    seeds = [rank * C + c for c in range(C)]
    if grid is not None:
        units, seeds, cpp = grid
        C = len(units)
    io = {"h2d": 0, "d2h": 0}

    if w["models"]:
        states = [loaders.init_random_weights(s) for s in range(w["models"])]     # random-init per BASELINE.json; untimed
        # edge capacity: a half-covered slab reaches ~130 neighbours per atom inside the 6 A skin list
        eng = engine.PainnEngine(states, pots["offset_data"] if name == "sto_painn" else None, edges_per_atom=192)
        to_species = lambda zz: zz
        if not os.environ.get("VSSR_NO_FILTER_MEMO"):
            # radial-filter memo for the frozen bulk (one-time, untimed); the relaxation holds those atoms with
            # FixAtoms, so their (discarded) force rows are not computed either
            eng.set_framework(pos, cell, pbc, fixed, constrained_forces=not os.environ.get("VSSR_FULL_GRAD"))

        def relax_batch(b, zh):
            return eng.relax(b, relax_steps=steps_relax, fmax=0.01, z_host=zh, want_std=False)
    else:
        if name == "gan_tersoff":
            tmap = {31: 0, 7: 1}
            eng = engine.ClassicalEngine(engine.POT_TERSOFF, engine.tersoff_param_table(pots["GaN.tersoff"], ["Ga", "N"]), 2,
                                         n_max=64, max_nbr=32)
        else:
            tmap = {14: 0}
            eng = engine.ClassicalEngine(engine.POT_SW, engine.sw_param_table(), 1, n_max=160, max_nbr=40)
        lut = np.zeros(120, np.int32)
        for zz_, t_ in tmap.items():
            lut[zz_] = t_
        to_species = lambda zz: lut[zz]

        def relax_batch(b, zh):
            return eng.relax(b, relax_steps=steps_relax, fmax=0.01, check=False)

    statuses = []

    class Pending:
        """Asynchronous engine call: the 8 scalars per chain are copied to pinned host memory behind the
        relaxation on the same stream; result() waits for that copy only."""

        pool = {}      # pinned staging buffers are recycled: cudaHostAlloc costs milliseconds

        def __init__(self, out_dev):
            key = (tuple(out_dev.shape), out_dev.dtype)
            free = Pending.pool.setdefault(key, [])
            self.key = key
            self.host = free.pop() if free else torch.empty(out_dev.shape, dtype=out_dev.dtype, pin_memory=True)
            self.host.copy_(out_dev, non_blocking=True)
            self.done = torch.cuda.Event()
            self.done.record()

        def result(self):
            self.done.synchronize()
            out = self.host.numpy().copy()
            Pending.pool[self.key].append(self.host)
            return out

    # e2e driver: chain groups of MultiChainMC.pipeline.  The classical kernels are one latency-bound CTA per chain and
    # leave most of the GPU idle at 256 chains, so their groups run on separate CUDA streams: the relaxations of the
    # groups overlap each other AND the host-side proposal / Metropolis work (the engine holds no shared workspace).
    # The PaiNN engine fills the GPU with one batch and owns one workspace: one group, default stream.
    n_groups = max(1, args.groups) if (args.groups or w["models"]) else 4
    side_streams = [torch.cuda.Stream() for _ in range(n_groups)] if (n_groups > 1 and not w["models"]) else []
    calls = [0]

    def relax_fn(pos_l, num_l, fix_l):
        stream = side_streams[calls[0] % n_groups] if side_streams else None
        calls[0] += 1
        with (torch.cuda.stream(stream) if stream is not None else contextlib.nullcontext()):
            b = engine.Batch.from_arrays(pos_l, [to_species(zz) for zz in num_l], [cell] * len(pos_l), [pbc] * len(pos_l), fix_l)
            r = relax_batch(b, np.concatenate(num_l))
            statuses.append(r["status"])              # checked once after the timed regions (no sync here)
            io["h2d"] += b.h2d_bytes() + 8 * b.n_struct
            io["d2h"] += r["out"].numel() * 8
            return Pending(r["out"])

    def join_side_streams():
        for s_ in side_streams:
            torch.cuda.current_stream().wait_stream(s_)

    def check_statuses():
        bits = 0
        for st_ in statuses:
            bits |= int(st_.item())
        statuses.clear()
        if bits:
            what = [n_ for b_, n_ in ((1, "edge capacity (edges_per_atom)"), (2, "neighbour slots (max_nbr)"), (4, "atoms per structure (n_max)")) if bits & b_]
            raise RuntimeError(f"{name}: a relaxation reported a device-side overflow of {', '.join(what)}: the run is invalid")

    drv = build_driver(name, relax_fn, seeds, units)
    if w["canonical"]:
        drv.prepare_canonical()
    drv._ensure_prev(drv.chains)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ------------------------------------------------------------ burn-in (untimed) + e2e: public API, host buffers
    pipe = drv.pipeline(n_groups=n_groups)
    t_burn = time.perf_counter()
    for k in range(burn_in):
        pipe.advance()
        if k % 50 == 49:
            check_statuses()
    torch.cuda.synchronize()
    t_burn = time.perf_counter() - t_burn
    n_ads = np.array([c.num_adsorbates for c in drv.chains])
    n_atoms = np.array([len(c) for c in drv.chains])
    coverage = {"burn_in_mc_steps": burn_in, "burn_in_seconds": round(t_burn, 2),
                "adsorbates_per_chain": {"mean": float(n_ads.mean()), "min": int(n_ads.min()), "max": int(n_ads.max())},
                "atoms_per_structure": {"mean": float(n_atoms.mean()), "max": int(n_atoms.max())},
                "accept_rate_burn_in": float(np.mean([d[0] for ch in drv.decisions for d in ch])) if burn_in else None}
    for _ in range(warmup):
        pipe.advance()
    sampler = ClockSampler(local, enabled=(rank == 0 and headline))   # one poller per job: N pollers contend for the driver
    barrier()
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = int(lib.vssr_launch_count())
    io["h2d"] = io["d2h"] = 0
    ev0.record()
    for _ in range(steps):
        pipe.advance()
    pipe.drain()                # (in-flight iterations of the pipelined groups belong to the timed steps)
    join_side_streams()
    ev1.record()
    barrier()
    e2e_ms = ev0.elapsed_time(ev1)
    e2e_launches = int(lib.vssr_launch_count()) - l0
    io = {k: v // steps for k, v in io.items()}
    pipe.drain()
    check_statuses()
    n_ads_end = float(np.mean([c.num_adsorbates for c in drv.chains]))

    # ------------------------------------------------------------ device-resident: relax call only
    # stage the proposal batches (current chain states + one fresh proposal each) in HBM beforehand
    staged, staged_host, atoms_total = [], [], 0
    for k in range(steps + warmup):
        pl, nl, fl = [], [], []
        for c in drv.chains:
            snap = c.snapshot()
            c.apply(c.propose_switch() if w["canonical"] else c.propose_change(w["adsorbates"]))
            p, zz = c.arrays()
            c.restore(snap)
            pl.append(p); nl.append(zz)
            fl.append(np.concatenate([fixed, np.zeros(len(zz) - len(fixed), bool)]))
        b = engine.Batch.from_arrays(pl, [to_species(zz) for zz in nl], [cell] * C, [pbc] * C, fl)
        staged.append((b, np.concatenate(nl)))
        if k >= warmup:
            atoms_total += b.n_atoms
            if len(staged_host) < 2:
                staged_host.append((pl, nl, fl))
    for b, zh in staged[:warmup]:
        statuses.append(relax_batch(b, zh)["status"])
    barrier()
    l0, g0 = int(lib.vssr_launch_count()), int(lib.vssr_graph_launch_count())
    ev0.record()
    for b, zh in staged[warmup:]:
        statuses.append(relax_batch(b, zh)["status"])
    ev1.record()
    barrier()
    dev_ms = ev0.elapsed_time(ev1)
    launches = int(lib.vssr_launch_count()) - l0                 # kernels executed (graph replays count their nodes)
    graph_launches = int(lib.vssr_graph_launch_count()) - g0      # of which arrived through cudaGraphLaunch calls
    clocks = sampler.stop()
    check_statuses()

    # ------------------------------------------------------------ per-kernel-class profile (two relaxations again)
    ncls = int(lib.vssr_kernel_class_count())
    ms = np.zeros(ncls); cnt = np.zeros(ncls, np.int64)
    fresh = [engine.Batch.from_arrays(pl, [to_species(zz) for zz in nl], [cell] * C, [pbc] * C, fl) for pl, nl, fl in staged_host]
    torch.cuda.synchronize()
    lib.vssr_profile_enable(1)
    edge_stats = []
    for b, (_, zh) in zip(fresh, staged[warmup:][:2]):
        relax_batch(b, zh)
        if w["models"]:
            edge_stats.append(eng.last_relax_edge_stats(b))
    lib.vssr_profile_collect(ms.ctypes.data, cnt.ctypes.data, ncls)
    lib.vssr_profile_enable(0)
    names = ["nbr", "edge_geometry", "gemm_tcgen05_3xtf32", "message_fwd", "message_bwd", "elementwise", "readout",
             "ensemble_stats", "fire", "classical_relax", "message_fwd_memo", "message_bwd_memo"]
    breakdown = {names[k]: {"ms": round(float(ms[k]), 3), "launches": int(cnt[k])} for k in range(ncls) if cnt[k]}
    a_prof = sum(b.n_atoms for b in fresh)

    # max over ranks
    t = torch.tensor([e2e_ms, dev_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms, dev_ms = t.tolist()
    c_tot = torch.tensor([float(C)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(c_tot, op=dist.ReduceOp.SUM)
    total_props = c_tot.item() * steps
    value = total_props / (dev_ms * 1e-3)
    e2e = total_props / (e2e_ms * 1e-3)

    peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else None
    extra = {}
    if w["models"]:
        roof = painn_roofline(breakdown, edge_stats, a_prof, steps_relax + 1, w["models"], peaks)
        extra["painn_atom_model_evals_per_sec"] = atoms_total * (steps_relax + 1) * w["models"] * world / (dev_ms * 1e-3)
    else:
        roof = classical_roofline(breakdown, a_prof, steps_relax + 1, peaks, name, proposals=C * len(fresh))
        # saturation point (SURVEY.md 8d caveat): 256 chains = 256 CTAs occupy a fraction of the 148 SMs x 2-4 resident
        # CTAs; the same kernel on 65,536 chains (the staged batch replicated) shows what the GPU sustains
        reps = 65536 // C
        pl, nl, fl = staged_host[0]
        big = engine.Batch.from_arrays(pl * reps, [to_species(zz) for zz in nl] * reps, [cell] * (C * reps), [pbc] * (C * reps), fl * reps)
        pos0 = big.pos.clone()
        relax_batch(big, None)               # warm-up (the relaxation works in place)
        big.pos.copy_(pos0)
        torch.cuda.synchronize()
        ev0.record()
        st_big = relax_batch(big, None)["status"]
        ev1.record()
        torch.cuda.synchronize()
        extra["saturation"] = {"chains_per_gpu": C * reps, "value": C * reps / (ev0.elapsed_time(ev1) * 1e-3), "unit": "proposals/s",
                               "ms": ev0.elapsed_time(ev1), "status_bits": int(st_big.item()),
                               "note": "device-resident, one relax call, the burnt-in batch replicated; per-GPU figure"}
    cfg_extra = {"coverage": {**coverage, "mean_adsorbates_after_timed_region": n_ads_end},
                 "timed_region_s": {"device": round(dev_ms * 1e-3, 3), "e2e": round(e2e_ms * 1e-3, 3)}}
    if grid is not None:
        cfg_extra["grid"] = {"pH": PH_VALUES, "U_V": U_VALUES, "chains_per_point": cpp, "units_this_rank": C,
                             "sharding": "parallel.shard_grid: (pH, U, chain) units round-robin over ranks"}
    out = {
        "metric": "relaxed_proposals_per_sec", "value": value, "unit": "proposals/s", "n_gpus": world,
        "steps": steps, "warmup": warmup, "ms_per_step": dev_ms / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32" if w["models"] else "f64", "data": "synthetic",
        "config": workload_config(name, args, C, world, cfg_extra),
        "e2e": {"value": e2e, "unit": "proposals/s", "h2d_bytes_per_step": io["h2d"], "d2h_bytes_per_step": io["d2h"],
                "ms_per_step": e2e_ms / steps},
        "gpu_launches": launches, "gpu_launches_e2e": e2e_launches, "graph_launches": graph_launches,
        "roofline": roof, "kernel_breakdown_ms": breakdown, "clocks": clocks, **extra,
    }
    if name in REFERENCE_PUBLISHED:
        out["reference_published_single_chain"] = REFERENCE_PUBLISHED[name]
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        occ = typical_occupancy(drv)
        n_cpu = args.cpu_baseline_proposals if (headline and args.cpu_baseline_proposals) else w["cpu_props"]
        n, dt, ev = oracle_proposals(name, n_cpu, threads, burn_occ=occ)
        out["cpu_baseline"] = {"value": n / dt, "unit": "proposals/s", "cores": threads, "kind": "port",
                               "sample": f"{n} single-chain relaxed proposals of the same workload ({dt:.1f} s) starting from the "
                                         f"burnt-in population's median coverage ({len(occ)} adsorbates), oracle port of the "
                                         "reference path on the host cores",
                               "painn_atom_model_evals_per_sec": ev / dt if ev else None}
        if w["models"] and headline and not args.no_torch_gpu_baseline:
            n, dt, ev = oracle_proposals(name, 2, threads, device="cuda", burn_occ=occ)       # warm-up (cuDNN/cuBLAS init)
            n, dt, ev = oracle_proposals(name, args.torch_gpu_proposals, threads, device="cuda", burn_occ=occ)
            out["reference_equivalent_gpu"] = {
                "value": n / dt, "unit": "proposals/s", "kind": "torch-eager-on-B200",
                "sample": f"{n} single-chain relaxed proposals ({dt:.1f} s): the oracle's torch fp32 PaiNN ensemble + autograd forces "
                          "run eagerly on this GPU, host FIRE, one chain -- how the reference drives its GPU (BASELINE.md section 2)",
                "painn_atom_model_evals_per_sec": ev / dt if ev else None}
    return out


def painn_roofline(breakdown, edge_stats, a_prof, evals, M, peaks):
    """Per kernel class: measured ms (CUDA-event pairs, 2 relaxations), ALGORITHMIC FLOPs and EXECUTED
    work, each against the peak of the pipe the kernels run on.  No class is charged for work it skips."""
    bf16 = peaks["bf16_tflops_sustained"] if peaks else 1400.0
    hbm = peaks["hbm_gbs"] if peaks else 6650.0
    tf32 = bf16 / 2
    src = "measured (MEASURED_PEAKS.json: sustained bf16 / 2 = TF32 dense; hbm_gbs)" if peaks else "fallback (1.4 PF / 2; 6.65 TB/s)"
    d = sum(s["direct_edges"] for s in edge_stats)
    m = sum(s["memo_edges"] for s in edge_stats)
    df = sum(s["direct_edges_frozen_receiver"] for s in edge_stats)
    constrained = "message_bwd_memo" in breakdown and not os.environ.get("VSSR_FULL_GRAD")
    ms = lambda k: breakdown.get(k, {"ms": 0.0})["ms"]
    slot = lambda n_edges, slots: n_edges * slots * 64 * 4 * 3 * M     # 64 lanes x (2 features x 2 flop) x 3 layers x models
    # (evals-1) constrained evaluations + 1 full-gradient evaluation per relaxation
    ce, fe = (evals - 1, 1) if constrained else (0, evals)
    ex_fwd = evals * (slot(d, MSG_SLOTS["direct_fwd"]) + slot(m, MSG_SLOTS["memo_fwd"]))
    ex_bwd = (ce * (slot(d - df, MSG_SLOTS["direct_bwd"]) + slot(df, MSG_SLOTS["direct_bwd_frozen_receiver"])
                    + slot(m, MSG_SLOTS["memo_bwd_state"]) * 2 / 3)           # constrained: nothing at layer 0
              + fe * (slot(d, MSG_SLOTS["direct_bwd"]) + slot(m, MSG_SLOTS["memo_bwd_full"])))
    classes = {}

    def add(key, t_ms, algo_flop, exec_flop, peak, pipe, bytes_algo=None):
        if t_ms <= 0:
            return
        c = {"ms": round(t_ms, 3), "pipe": pipe}
        if algo_flop is not None:
            c["algorithmic_tflops"] = algo_flop / (t_ms * 1e-3) / 1e12
            c["frac_algorithmic_of_tf32_tensor_peak"] = c["algorithmic_tflops"] / tf32
        if exec_flop is not None:
            c["executed_tflops"] = exec_flop / (t_ms * 1e-3) / 1e12
            c["frac_executed_of_pipe_peak"] = c["executed_tflops"] / peak
            c["pipe_peak_tflops"] = peak
        if bytes_algo is not None:
            c["achieved_gbs"] = bytes_algo / (t_ms * 1e-3) / 1e9
            c["frac_of_hbm_peak"] = c["achieved_gbs"] / hbm
        classes[key] = c

    E = d + m
    gemm_algo = 2 * GEMM_FLOP_PER_ATOM * M * a_prof * evals                 # forward + backward
    add("gemm", ms("gemm_tcgen05_3xtf32"), gemm_algo, 3 * gemm_algo, tf32, "tensor (tcgen05 kind::tf32, 3 MMAs per algorithmic MAC)")
    add("message_fwd", ms("message_fwd") + ms("message_fwd_memo"), FILTER_FLOP_PER_EDGE * E * M * evals, ex_fwd, FFMA2_PEAK,
        "fp32 FMA (packed FFMA2)")
    add("message_bwd", ms("message_bwd") + ms("message_bwd_memo"), FILTER_FLOP_PER_EDGE * E * M * evals, ex_bwd, FFMA2_PEAK,
        "fp32 FMA (packed FFMA2)")
    # element-wise glue: ~ (2..12 floats per atom-feature) per kernel; algorithmic bytes of the un-fused dataflow
    ew_bytes = 4.0 * 128 * M * a_prof * evals * 3 * (8 + 16 + 19 + 9)        # nrm, update_fwd, update_bwd, nrm_bwd per layer
    add("elementwise", ms("elementwise"), None, None, None, "HBM", bytes_algo=ew_bytes)
    if not classes:
        return None
    dom = max(classes, key=lambda k: classes[k]["ms"])
    c = classes[dom]
    traffic, traffic_src, per_eval = None, None, None
    caps = sorted((ROOT / "profiles").glob("r*_ncu_traffic.json"))
    if caps:
        cap = json.loads(caps[-1].read_text())
        key = {"gemm": "gemm_tcgen05_3xtf32"}.get(dom, dom)
        if key in cap.get("classes", {}):
            traffic = cap["classes"][key]["dram_bytes_per_launch"]
            traffic_src = f"profiles/{caps[-1].name} (mean dram__bytes_read+write per launch, {cap['classes'][key]['launches_captured']} launches)"
        per_eval = cap.get("dram_bytes_per_evaluation")
    is_tensor = dom == "gemm"
    total_ms = sum(v["ms"] for v in breakdown.values())
    algo_all = (gemm_algo + 2 * FILTER_FLOP_PER_EDGE * E * M * evals)
    return {"bound": "tensor" if is_tensor else "fma", "kernel": dom,
            "achieved": c["algorithmic_tflops"] if is_tensor else c["executed_tflops"],
            "peak": tf32 if is_tensor else FFMA2_PEAK, "unit": "TFLOP/s",
            "frac": c["frac_algorithmic_of_tf32_tensor_peak"] if is_tensor else c["frac_executed_of_pipe_peak"],
            "frac_algorithmic_of_tf32_tensor_peak": c["frac_algorithmic_of_tf32_tensor_peak"],
            "traffic": traffic, "traffic_source": traffic_src, "peak_source": src + f"; FFMA2 {FFMA2_PEAK} TFLOP/s measured (profiles/microbench/ffma2.cu)",
            "classes": classes,
            "whole_step": {"algorithmic_tflops": algo_all / (total_ms * 1e-3) / 1e12,
                           "frac_of_tf32_tensor_peak": algo_all / (total_ms * 1e-3) / 1e12 / tf32,
                           "class_ms_sum": round(total_ms, 2)},
            "edges_per_evaluation": {"direct": d // max(len(edge_stats), 1), "memoised": m // max(len(edge_stats), 1),
                                     "direct_with_frozen_receiver": df // max(len(edge_stats), 1),
                                     "canonical_structures": edge_stats[0]["canonical_structures"] if edge_stats else None,
                                     "structures": edge_stats[0]["structures"] if edge_stats else None},
            "dram_bytes_per_evaluation": {"measured_ncu": per_eval, "fused_ideal": 320.0 * a_prof / max(len(edge_stats), 1),
                                          "note": "SURVEY.md 8d: 320 B per atom-evaluation if the whole model were one fused kernel"},
            "note": "message kernels run on the fp32 FMA pipe: `frac` = executed packed-FMA issue slots / measured FFMA2 peak; the "
                    "north star's tensor-core bound for the K=20 filter contraction is reported as frac_algorithmic_of_tf32_tensor_peak"
                    if not is_tensor else "3xTF32 on tcgen05: `frac` = algorithmic FLOPs / TF32 peak (x3 in issued tensor MACs)"}


def classical_roofline(breakdown, a_prof, evals, peaks, name=None, proposals=0):
    """Persistent one-CTA-per-chain kernel: the slab never leaves shared memory between FIRE steps, so HBM traffic is ~0
    and the kernel is FP64-pipe / latency bound.  Reported against the FP64 pipe: executed FP64 flops per relaxed
    proposal (ncu instruction counts on this bench's own launches, profiles/round2_classical_fp64.json) x the proposals
    of the profiled launches / their time; the algorithmic HBM figure of SURVEY.md 8d (53 B per atom-evaluation if
    the kernel streamed) is kept beside it."""
    hbm = peaks["hbm_gbs"] if peaks else 6650.0
    k = "classical_relax"
    if k not in breakdown or breakdown[k]["ms"] <= 0:
        return None
    t = breakdown[k]["ms"] * 1e-3
    streamed = 53.0 * a_prof * evals / t / 1e9
    roof = {"bound": "fp64", "kernel": k, "achieved": None, "peak": FP64_PEAK, "unit": "TFLOP/s", "frac": None, "traffic": None,
            "peak_source": "nominal B200 FP64 (148 SMs x 64 DFMA/clk x 2 x 1.965 GHz); MEASURED_PEAKS.json has no FP64 entry",
            "atom_evals_per_sec_upper_bound": a_prof * evals / t,
            "hbm_if_streamed": {"achieved_gbs": streamed, "peak_gbs": hbm, "frac": streamed / hbm,
                                "note": "SURVEY.md 8d: 53 B per atom-evaluation; the kernel keeps the relaxation in shared "
                                        "memory, so this only shows that it is not bandwidth-bound"},
            "note": "one 256-thread CTA per chain (thread per directed atom pair), whole relaxation (<= 100 dependent FIRE "
                    "steps, converged chains stop early) inside the SM: latency-bound per chain, the chain count is the "
                    "parallelism (see `saturation`)"}
    f = ROOT / "profiles" / "round2_classical_fp64.json"
    if f.exists() and name:
        d = json.loads(f.read_text()).get(name)
        if d:
            roof["achieved"] = d["fp64_flop_per_proposal"] * proposals / t / 1e12
            roof["frac"] = roof["achieved"] / FP64_PEAK
            roof["fp64_flop_per_proposal"] = d["fp64_flop_per_proposal"]
            roof["flop_source"] = "profiles/round2_classical_fp64.json: " + d["source"]
    return roof


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None, choices=list(WORKLOADS))
    ap.add_argument("--chains-per-gpu", type=int, default=0)
    ap.add_argument("--burn-in", type=int, default=-1, help="untimed MC steps before the timed region (default: per workload)")
    ap.add_argument("--groups", type=int, default=0,
                    help="interleaved chain groups of the e2e driver (MultiChainMC.pipeline); default: 1 for PaiNN (the host part "
                         "of a step is ~3 ms and half-size batches cost the GPU more than that), 4 on separate streams for the "
                         "classical potentials")
    ap.add_argument("--cpu-baseline-proposals", type=int, default=0)
    ap.add_argument("--torch-gpu-proposals", type=int, default=6)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-torch-gpu-baseline", action="store_true")
    ap.add_argument("--no-extra-workloads", action="store_true")
    args = ap.parse_args()
    explicit = args.workload is not None
    args.workload = args.workload or "sto_painn"
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from surface_sampling_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    torch.set_num_threads(max(1, (os.cpu_count() or 1) // max(world, 1) // 2))   # host threads are for launching, not for BLAS
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = {"lib": _lib.load(), "world": world, "rank": rank, "local": local}
    name = args.workload
    burn = args.burn_in if args.burn_in >= 0 else WORKLOADS[name]["burn_in"]
    out = run_workload(name, args, ctx, args.steps, args.warmup, burn, headline=True)
    if not explicit and not args.no_extra_workloads:
        # the other BASELINE configs, short runs (their own value / e2e / roofline / cpu_baseline)
        out["workloads"] = []
        for other in ("gan_tersoff", "si_sw", "sto_pourbaix"):
            try:
                o = run_workload(other, args, ctx, min(args.steps, 10), 3, WORKLOADS[other]["burn_in"], headline=False)
            except Exception as exc:      # a side workload must not cost the headline line
                out["workloads"].append({"config": {"workload": WORKLOADS[other]["desc"]}, "error": f"{type(exc).__name__}: {exc}"})
                continue
            out["workloads"].append({k: o[k] for k in ("metric", "value", "unit", "n_gpus", "steps", "ms_per_step", "dtype", "config",
                                                       "e2e", "gpu_launches", "roofline", "kernel_breakdown_ms", "reference_published_single_chain", "saturation") if k in o}
                                    | ({"cpu_baseline": o["cpu_baseline"]} if "cpu_baseline" in o else {}))
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
