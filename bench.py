#!/usr/bin/env python
"""bench.py — relaxed VSSR-MC proposals/s (headline: SrTiO3(001) 2x2, PaiNN 3-model ensemble).

A "step" is one MC iteration of every chain on this GPU: propose (host, reference RNG order) ->
ideal-site structure -> H2D -> neighbour list + FIRE relaxation with ensemble force evaluations
(GPU, no host round trip) -> 8 scalars per chain D2H -> Metropolis accept/reject.

  value   = relaxed proposals/s with the step's inputs already resident in HBM (relax call only)
  e2e     = the same metric through the public API (MultiChainMC.step) with HOST buffers:
            pinned H2D of positions/species/masks and D2H of the result inside the timed region
  roofline= dominant kernel class, timed live with CUDA-event pairs on the launching stream
  cpu_baseline = the CPU oracle (torch fp32 PaiNN + autograd forces + numpy FIRE) on the host cores

`--impl reference` times the reference path's CPU restatement (the oracle: the reference's own
stack — ASE/NFF/LAMMPS — is not installable here) on the same config.
Weak scaling: --chains-per-gpu chains on every rank (128 -> 1024 chains on 8 GPUs, BASELINE
config 4); chains never interact, NCCL only gathers per-chain scalars.

Other BASELINE configs (parity-test cases, not the headline line): --workload gan_tersoff
(config 2: GaN(0001) Tersoff, canonical 12 Ga, relax_steps 100) and --workload si_sw (config 3:
Si(111) 5x5, SW, relax per proposal).
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
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
GOLD = ROOT / "tests" / "golden"

CHEM_POTS = {"Sr": -2, "Ti": 0, "O": 0}
FREE = [7, 8, 22, 23, 37, 38, 52, 53]   # surface_depth=1 (tutorials/SrTiO3_001.ipynb cell 7 log)
FFMA2_PEAK = 65.8   # TFLOP/s, packed fp32x2 FMA measured on this pool's B200 (profiles/microbench/ffma2.cu)

WORKLOADS = {
    "sto_painn": dict(slab="SrTiO3_001_2x2", n_sites=64, adsorbates=["Sr", "Ti", "O"], relax_steps=20, canonical=False,
                      num_ads=0, height=1.5, chains=128,
                      desc="SrTiO3(001) 2x2 VSSR-MC, PaiNN 3-model ensemble (random-init weights seeds 0,1,2), semigrand "
                           "Sr/Ti/O on 64 virtual sites, FIRE relax_steps=20 fmax=0.01 (BASELINE.json configs[3])"),
    "gan_tersoff": dict(slab="GaN_0001_3x3", n_sites=107, adsorbates=["Ga"], relax_steps=100, canonical=True, num_ads=12,
                        height=1.8, chains=256,
                        desc="GaN(0001) 3x3 VSSR-MC, Tersoff (Nord 2003), canonical 12 Ga adatoms on 107 virtual sites, "
                             "FIRE relax_steps=100 fmax=0.01, bulk ids<=36 frozen (BASELINE.json configs[1])"),
    "si_sw": dict(slab="Si_111_5x5", n_sites=100, adsorbates=["Si"], relax_steps=100, canonical=False, num_ads=0,
                  height=2.0, chains=256,
                  desc="Si(111) 5x5 VSSR-MC, Stillinger-Weber (SW-1985 literature parameters, parity unpinned), semigrand "
                       "Si on 100 virtual sites, FIRE relax_steps=100 fmax=0.01, ids<=75 frozen (BASELINE.json configs[2])"),
}


def load_workload(name):
    w = WORKLOADS[name]
    z = np.load(GOLD / "structures.npz")
    pots = json.loads((GOLD / "potentials.json").read_text())
    n = w["slab"]
    pos, num, cell, pbc = z[f"{n}/positions"], z[f"{n}/numbers"], z[f"{n}/cell"], z[f"{n}/pbc"]
    if name == "sto_painn":
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

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
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
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
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


def make_oracle_relax_fn(name, evals):
    """Single-chain CPU restatement of the hot path for `name` (oracle physics + oracle FIRE)."""
    import torch
    from oracle import classical as ocl
    from oracle import relax as orelax
    from oracle.painn import EnsembleOracle, init_random_weights

    pos0, num0, cell, pbc, fixed0, pots = load_workload(name)
    w = WORKLOADS[name]
    if name == "sto_painn":
        ens = EnsembleOracle([init_random_weights(s) for s in (0, 1, 2)], pots["offset_data"], dtype=torch.float32)

        def energy_forces_factory(p, zz):
            nb = ens.build_nbrs(p, cell, pbc)

            def calc(x):
                r = ens.calculate(x, zz, cell, pbc, nb)
                evals[0] += 3 * len(zz)
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


def surface_energy_fn(name, pots):
    if name != "sto_painn":
        return lambda e, sym: e           # LAMMPSSurfCalc: surface energy = potential energy
    from surface_sampling_b200.calculators import surface_energy_from
    od = pots["offset_data"]
    return lambda e, sym: surface_energy_from(e, sym, od, CHEM_POTS)


def build_driver(name, relax_fn, seeds):
    from surface_sampling_b200 import mc
    pos, num, cell, pbc, fixed, pots = load_workload(name)
    w = WORKLOADS[name]
    sites = mc.make_site_grid(pos, cell, w["n_sites"], w["height"])
    return mc.MultiChainMC(num, pos, fixed, sites, w["adsorbates"], relax_fn, surface_energy_fn(name, pots), seeds,
                           canonical=w["canonical"], num_ads_atoms=w["num_ads"])


def cpu_oracle_proposals(name: str, n_proposals: int, threads: int, seed0: int = 0):
    """Reference-path CPU restatement: single chain.  Returns (proposals, seconds, atom_model_evals)."""
    import torch
    torch.set_num_threads(threads)
    evals = [0]
    drv = build_driver(name, make_oracle_relax_fn(name, evals), [seed0])
    if WORKLOADS[name]["canonical"]:
        drv.prepare_canonical()
    drv._ensure_prev(drv.chains)   # the initial-state energy is not a proposal
    drv.n_relaxed, evals[0] = 0, 0
    t0 = time.perf_counter()
    for _ in range(n_proposals):
        drv.step()
    dt = time.perf_counter() - t0
    return drv.n_relaxed, dt, evals[0]


def workload_config(args, chains):
    return {"workload": WORKLOADS[args.workload]["desc"], "chains_per_gpu": chains,
            "l2": "per-evaluation working set (activations ~48 KB/atom/model, >1 GB) exceeds the 126 MB L2; no explicit flush"
                  if args.workload == "sto_painn" else "whole relaxation is shared-memory resident; L2 is not on the path",
            "parallelism": f"chains sharded, {args.gpus} rank(s), no data-path collective",
            "e2e_driver": f"MultiChainMC.pipeline, {max(1, args.groups)} chain group(s) per GPU; value = one batch of all chains per step",
            "engine_options": ({"filter_memo": not os.environ.get("VSSR_NO_FILTER_MEMO"),
                                "constrained_gradients": not (os.environ.get("VSSR_FULL_GRAD") or os.environ.get("VSSR_NO_FILTER_MEMO")),
                                "note": "every evaluation runs the full 3-layer forward and backward of all 3 models; the memo holds the "
                                        "radial filter rows w(d), dw/dd of frozen-frozen pairs (functions of weights and the frozen "
                                        "geometry only), and dE/dx of FixAtoms atoms -- which the optimiser discards -- is not formed "
                                        "during FIRE steps; energies, positions and accept/reject are bit-identical to the plain mode "
                                        "(tests/test_gpu_painn.py); VSSR_FULL_GRAD=1 / VSSR_NO_FILTER_MEMO=1 switch them off"}
                               if args.workload == "sto_painn" else None)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    for _ in range(min(args.warmup, 1)):
        cpu_oracle_proposals(args.workload, 1, threads)
    n, dt, evals = 0, 0.0, 0
    for s in range(args.steps):
        a, b, c = cpu_oracle_proposals(args.workload, 1, threads, seed0=s)
        n, dt, evals = n + a, dt + b, evals + c
    val = n / dt
    print(json.dumps({
        "impl": "reference", "metric": "relaxed_proposals_per_sec", "value": val, "unit": "proposals/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if args.workload == "sto_painn" else "f64", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": val, "unit": "proposals/s", "cores": threads, "kind": "port",
                         "sample": f"{n} single-chain relaxed proposals (1 per step), oracle port of the reference path "
                                   "(torch-CPU physics + numpy FIRE); the reference stack itself is not installable here"},
        "e2e": {"value": val, "unit": "proposals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "painn_atom_model_evals_per_sec": evals / dt if evals else None,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="sto_painn", choices=list(WORKLOADS))
    ap.add_argument("--chains-per-gpu", type=int, default=0)
    ap.add_argument("--groups", type=int, default=1,
                    help="interleaved chain groups of the e2e driver (MultiChainMC.pipeline); 1 = plain lock step, which is "
                         "fastest here: the host part of a step is ~3 ms and half-size batches cost the GPU more than that")
    ap.add_argument("--cpu-baseline-proposals", type=int, default=16)   # ~15 s of host work
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from surface_sampling_b200 import _lib, engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    torch.set_num_threads(max(1, (os.cpu_count() or 1) // max(world, 1) // 2))   # host threads are for launching, not for BLAS
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = _lib.load()
    name = args.workload
    w = WORKLOADS[name]
    C = args.chains_per_gpu or w["chains"]
    steps_relax = w["relax_steps"]
    pos, num, cell, pbc, fixed, pots = load_workload(name)
    io = {"h2d": 0, "d2h": 0}

    if name == "sto_painn":
        from oracle.painn import init_random_weights   # weight INIT only (random-init per BASELINE.json); not timed
        eng = engine.PainnEngine([init_random_weights(s) for s in (0, 1, 2)], pots["offset_data"])
        e_cap = C * 72 * 96
        to_species = lambda zz: zz
        if not os.environ.get("VSSR_NO_FILTER_MEMO"):
            # radial-filter memo for the frozen bulk (one-time, untimed); the relaxation holds those atoms with
            # FixAtoms, so their (discarded) force rows are not computed either
            eng.set_framework(pos, cell, pbc, fixed, constrained_forces=not os.environ.get("VSSR_FULL_GRAD"))

        def relax_batch(b, zh):
            return eng.relax(b, relax_steps=steps_relax, fmax=0.01, z_host=zh, want_std=False, e_cap=e_cap)
    else:
        if name == "gan_tersoff":
            tmap = {31: 0, 7: 1}
            eng = engine.ClassicalEngine(engine.POT_TERSOFF, engine.tersoff_param_table(pots["GaN.tersoff"], ["Ga", "N"]), 2,
                                         n_max=64, max_nbr=24)
        else:
            tmap = {14: 0}
            eng = engine.ClassicalEngine(engine.POT_SW, engine.sw_param_table(), 1, n_max=128, max_nbr=32)
        to_species = lambda zz: np.array([tmap[int(q)] for q in zz], np.int32)

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

    def relax_fn(pos_l, num_l, fix_l):
        b = engine.Batch.from_arrays(pos_l, [to_species(zz) for zz in num_l], [cell] * len(pos_l), [pbc] * len(pos_l), fix_l)
        r = relax_batch(b, np.concatenate(num_l))
        statuses.append(r["status"].clone())          # checked once after the timed regions (no sync here)
        io["h2d"] += b.h2d_bytes() + 8 * b.n_struct
        io["d2h"] += r["out"].numel() * 8
        return Pending(r["out"])

    drv = build_driver(name, relax_fn, [rank * C + c for c in range(C)])
    if w["canonical"]:
        drv.prepare_canonical()
    drv._ensure_prev(drv.chains)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ------------------------------------------------------------ e2e: public API, host buffers
    # MultiChainMC.pipeline with `--groups` interleaved chain groups (with >1, the host applies Metropolis to /
    # proposes for one group while the GPU relaxes the other).  One step = one iteration of EVERY chain.
    pipe = drv.pipeline(n_groups=max(1, args.groups))
    for _ in range(args.warmup):
        pipe.advance()
    sampler = ClockSampler(local)
    if rank != 0:
        sampler.start = lambda: None      # one nvidia-smi poller per job (rank 0's GPU): N pollers contend for the driver
    barrier()
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = int(lib.vssr_launch_count())
    io["h2d"] = io["d2h"] = 0
    hostprof = None
    if os.environ.get("VSSR_BENCH_HOSTPROF"):      # where does the host spend the e2e step? (stderr, rank 0)
        import cProfile
        hostprof = cProfile.Profile()
        hostprof.enable()
    ev0.record()
    for _ in range(args.steps):
        pipe.advance()
    ev1.record()
    if hostprof is not None:
        hostprof.disable()
        if rank == 0:
            import pstats
            pstats.Stats(hostprof, stream=sys.stderr).sort_stats("tottime").print_stats(18)
    barrier()
    e2e_ms = ev0.elapsed_time(ev1)
    e2e_launches = int(lib.vssr_launch_count()) - l0
    io = {k: v // args.steps for k, v in io.items()}
    pipe.drain()
    if statuses and int(torch.stack(statuses).max().item()) != 0:
        raise RuntimeError("a relaxation reported a device-side overflow (edge capacity / neighbour slots): the run is invalid")

    # ------------------------------------------------------------ device-resident: relax call only
    # stage the proposal batches (current chain states + one fresh proposal each) in HBM beforehand
    staged, staged_host, atoms_total = [], [], 0
    for k in range(args.steps + args.warmup):
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
        if k >= args.warmup:
            atoms_total += b.n_atoms
            if len(staged_host) < 2:
                staged_host.append((pl, nl, fl))
    for b, zh in staged[:args.warmup]:
        relax_batch(b, zh)
    barrier()
    l0 = int(lib.vssr_launch_count())
    ev0.record()
    for b, zh in staged[args.warmup:]:
        relax_batch(b, zh)
    ev1.record()
    barrier()
    dev_ms = ev0.elapsed_time(ev1)
    launches = int(lib.vssr_launch_count()) - l0
    clocks = sampler.stop()

    # ------------------------------------------------------------ per-kernel-class profile (same region again)
    ncls = int(lib.vssr_kernel_class_count())
    ms = np.zeros(ncls); cnt = np.zeros(ncls, np.int64)
    lib.vssr_profile_enable(1)
    fresh = [engine.Batch.from_arrays(pl, [to_species(zz) for zz in nl], [cell] * C, [pbc] * C, fl) for pl, nl, fl in staged_host]
    for b, (_, zh) in zip(fresh, staged[args.warmup:][:2]):
        relax_batch(b, zh)
    lib.vssr_profile_collect(ms.ctypes.data, cnt.ctypes.data, ncls)
    lib.vssr_profile_enable(0)
    gemm_name = "gemm_fp32_ffma2" if os.environ.get("VSSR_GEMM", "tc").startswith("f") else "gemm_tcgen05_3xtf32"
    names = ["nbr", "edge_geometry", gemm_name, "message_fwd", "message_bwd", "elementwise", "readout",
             "ensemble_stats", "fire", "classical_relax", "message_fwd_memo", "message_bwd_memo"]
    breakdown = {names[k]: {"ms": round(float(ms[k]), 3), "launches": int(cnt[k])} for k in range(ncls) if cnt[k]}
    a_prof = sum(b.n_atoms for b in fresh)

    # max over ranks
    t = torch.tensor([e2e_ms, dev_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms, dev_ms = t.tolist()
    total_props = C * args.steps * world
    value = total_props / (dev_ms * 1e-3)
    e2e = total_props / (e2e_ms * 1e-3)

    # ------------------------------------------------------------ roofline of the dominant kernel class
    dom = max(breakdown, key=lambda k: breakdown[k]["ms"]) if breakdown else None
    if dom and dom.endswith("_memo"):
        dom = dom[:-5]
    peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else None
    roof = None
    extra = {}
    if name == "sto_painn":
        bf16 = peaks["bf16_tflops_sustained"] if peaks else 1400.0
        peak_src = "measured (MEASURED_PEAKS.json: sustained bf16 / 2 = TF32 dense)" if peaks else "fallback (1.4 PF / 2)"
        evals_per_prop = (steps_relax + 1) * 3      # model force evaluations per proposal (3-model ensemble)
        extra["painn_atom_model_evals_per_sec"] = atoms_total * evals_per_prop * world / (dev_ms * 1e-3)
        E_per_atom = 2504 / 60.0
        flops = {   # algorithmic FLOPs per model force-evaluation per atom (DESIGN.md section 2 / SURVEY.md 8d)
            gemm_name: 2 * 1491072.0,
            "message_fwd": 3 * E_per_atom * 2 * (3 * 20 * 128 + 12 * 128),
            "message_bwd": 3 * E_per_atom * 2 * (6 * 20 * 128 + 60 * 128),
        }

        def tflops(k):
            if k not in breakdown or breakdown[k]["ms"] <= 0:
                return None
            t_ms = breakdown[k]["ms"] + breakdown.get(k + "_memo", {"ms": 0.0})["ms"]   # both passes of a layer
            return flops[k] * 3 * a_prof * (steps_relax + 1) / (t_ms * 1e-3) / 1e12

        if dom in flops and tflops(dom):
            achieved = tflops(dom)
            is_gemm = dom == gemm_name
            # DRAM bytes per launch of the dominant kernel class from the committed ncu --set full capture
            traffic, traffic_src = None, None
            caps = sorted((ROOT / "profiles").glob("r*_ncu_traffic.json"))
            if caps:
                cap = json.loads(caps[-1].read_text())
                if dom in cap["classes"]:
                    traffic = cap["classes"][dom]["dram_bytes_per_launch"]
                    traffic_src = f"profiles/{caps[-1].name} (mean dram__bytes_read+write per launch, {cap['classes'][dom]['launches_captured']} launches)"
            roof = {"bound": "tensor", "kernel": dom, "achieved": achieved, "peak": bf16 / 2, "unit": "TFLOP/s",
                    "frac": achieved / (bf16 / 2), "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                    "achieved_algorithmic_tflops": {k: tflops(k) for k in flops},
                    "frac_of_measured_ffma2_peak": None if is_gemm else achieved / FFMA2_PEAK,
                    "note": ("3xTF32 on tcgen05: 3 tensor-core MACs per algorithmic MAC" if is_gemm else
                             "message passing runs on the fp32 FMA pipe (packed FFMA2, measured peak 65.8 TFLOP/s); the "
                             "tensor-pipe peak is quoted because the schema has no fp32-FMA bound")}
    else:
        hbm = peaks["hbm_gbs"] if peaks else 6650.0
        algo_bytes = 53.0 * a_prof * (steps_relax + 1)    # 53 B per atom-eval if it streamed
        if dom and breakdown[dom]["ms"] > 0:
            achieved = algo_bytes / (breakdown[dom]["ms"] * 1e-3) / 1e9
            roof = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
                    "traffic": None, "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback",
                    "note": "persistent one-CTA-per-chain kernel: the slab never leaves shared memory between FIRE steps, so "
                            "HBM traffic is ~0 and the kernel is FP64/latency-bound; the fraction only shows that"}
    out = {
        "metric": "relaxed_proposals_per_sec", "value": value, "unit": "proposals/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32" if name == "sto_painn" else "f64", "data": "synthetic",
        "config": workload_config(args, C),
        "e2e": {"value": e2e, "unit": "proposals/s", "h2d_bytes_per_step": io["h2d"], "d2h_bytes_per_step": io["d2h"],
                "ms_per_step": e2e_ms / args.steps},
        "gpu_launches": launches, "gpu_launches_e2e": e2e_launches,
        "roofline": roof, "kernel_breakdown_ms": breakdown, "clocks": clocks, **extra,
    }
    if rank == 0:
        if not args.no_cpu_baseline and world == 1:
            threads = os.cpu_count() or 1
            n, dt, ev = cpu_oracle_proposals(name, args.cpu_baseline_proposals, threads)
            out["cpu_baseline"] = {"value": n / dt, "unit": "proposals/s", "cores": threads, "kind": "port",
                                   "sample": f"{n} single-chain relaxed proposals of the same workload ({dt:.1f} s), "
                                             "oracle port of the reference path on the host cores",
                                   "painn_atom_model_evals_per_sec": ev / dt if ev else None}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
