"""Multi-chain driver vs the single-chain oracle loop under the same per-chain seeds: identical
accept/reject sequences and occupancies (synthetic energy function, CPU only).  Also the
world_size-2 gloo path of the per-sweep statistics gather."""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from oracle import mc as omc
from surface_sampling_b200 import mc
from surface_sampling_b200.engine import NUMBERS, SYMBOLS

ROOT = Path(__file__).resolve().parent.parent


def toy_energy(symbols, pos):
    """Deterministic, permutation-dependent-free toy 'relaxed surface energy'."""
    z = np.array([NUMBERS[s] for s in symbols], float)
    d = np.linalg.norm(pos[:, None, :] - pos[None, :, :], axis=-1) + np.eye(len(pos))
    return float(np.sum(np.triu(np.sqrt(z[:, None] * z[None, :]) * np.exp(-d) * np.cos(1.3 * d), 1)) - 0.05 * len(z))


def _driver(seeds, canonical=False, num_ads=0):
    numbers0 = [NUMBERS[s] for s in ["Sr", "Ti", "O", "O"]]
    pos0 = np.array([[0, 0, 0], [2, 0, 0], [0, 2, 0], [2, 2, 0.5]], float)
    sites = np.array([[x, y, 1.5] for x in (0.0, 1.0, 2.0) for y in (0.0, 1.0, 2.0)])

    def relax_fn(pos_l, num_l, fix_l):
        out = np.zeros((len(pos_l), 8))
        for k, (p, z) in enumerate(zip(pos_l, num_l)):
            out[k, 2] = out[k, 0] = toy_energy([SYMBOLS[int(q)] for q in z], p)
        return out

    drv = mc.MultiChainMC(numbers0, pos0, np.ones(4, bool), sites, ["Sr", "O"], relax_fn, lambda e, sym: e, seeds,
                          canonical=canonical, num_ads_atoms=num_ads)
    return drv, (["Sr", "Ti", "O", "O"], pos0, sites)


@pytest.mark.parametrize("canonical", [False, True])
def test_decisions_match_single_chain_oracle(canonical):
    seeds = [0, 1, 7, 11, 12345]
    drv, (sym0, pos0, sites) = _driver(seeds, canonical, 3 if canonical else 0)
    res = drv.run(total_sweeps=3, sweep_size=8, start_temp=0.5, perform_annealing=True, alpha=0.9)
    for k, s in enumerate(seeds):
        o = omc.run_chain(s, sym0, pos0, sites, ["Sr", "O"], toy_energy, 3, 8, start_temp=0.5, alpha=0.9,
                          canonical=canonical, num_ads_atoms=3 if canonical else 0)
        assert [d[0] for d in drv.decisions[k]] == [d[0] for d in o["decisions"]]
        assert [d[3] for d in drv.decisions[k]] == [d[3] for d in o["decisions"]]       # same uniforms
        assert np.allclose([d[1] for d in drv.decisions[k]], [d[1] for d in o["decisions"]], rtol=0, atol=0)
        assert list(drv.chains[k].occ) == o["final"].occ
        assert np.array_equal(res["energy_hist"][k], o["energy_hist"])
        assert np.array_equal(res["frac_accept_hist"][k], o["frac_accept_hist"])
        assert np.array_equal(res["adsorption_count_hist"][k], o["ads_hist"])


def test_overflowing_boltzmann_factor_accepts():
    drv, _ = _driver([3])
    drv.temp = 1e-320   # exp(+huge) -> inf -> accept (criterion.py:162-165 never catches numpy overflow)
    c = drv.chains[0]
    c.results["surface_energy"] = 1e6
    assert drv.step() == [True]


def _free_port() -> str:
    """A port the kernel just handed out: back-to-back runs of the suite must not trip over a socket in TIME_WAIT."""
    import socket
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as sk:
        sk.bind(("127.0.0.1", 0))
        return str(sk.getsockname()[1])


@pytest.mark.timeout(120)
def test_sharded_statistics_gather_gloo_world2(tmp_path):
    """Chains shard across ranks with no data-path collective; only per-sweep scalars are gathered."""
    script = ROOT / "tests" / "_gloo_worker.py"
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=_free_port(), PYTHONPATH=str(ROOT))
    procs = [subprocess.Popen([sys.executable, str(script), str(r), "2", str(tmp_path)], env=env) for r in range(2)]
    assert all(p.wait(timeout=100) == 0 for p in procs)
    got = np.load(tmp_path / "gathered.npy")
    drv, _ = _driver(list(range(6)))
    ref = drv.run(total_sweeps=2, sweep_size=4, start_temp=0.5, perform_annealing=False)
    assert np.array_equal(got[0], ref["energy_hist"]) and np.array_equal(got[2], ref["adsorption_count_hist"])


def _grid_driver(rank, world):
    """Toy slab, Pourbaix grand potential per (pH, U) unit exactly as bench.py wires BASELINE config 5."""
    import bench
    mine, seeds, cpp = bench.grid_units("sto_pourbaix", 2 * len(bench.PH_VALUES) * len(bench.U_VALUES) // world, rank, world)
    n_total = len(bench.PH_VALUES) * len(bench.U_VALUES) * cpp
    base, _ = _driver([0])
    drv = mc.MultiChainMC([NUMBERS[q] for q in ["Sr", "Ti", "O", "O"]], base.chains[0]._pos0, np.ones(4, bool),
                          base.chains[0].ads_coords, ["Sr", "O", "HO"], base.relax_fn,
                          bench.surface_energy_fns("sto_pourbaix", None, mine), seeds)
    return drv, n_total


@pytest.mark.timeout(180)
def test_pourbaix_grid_shards_over_ranks_gloo_world2(tmp_path):
    """BASELINE config 5: 8 pH x 7 U grid points x chains as independent (pH, U, chain) units, round-robin over ranks
    (parallel.shard_grid); every unit carries its own grand-potential scalar; only per-sweep scalars are gathered.
    Two gloo ranks reproduce the single-process run unit by unit."""
    script = ROOT / "tests" / "_gloo_worker.py"
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=_free_port(), PYTHONPATH=str(ROOT))
    procs = [subprocess.Popen([sys.executable, str(script), str(r), "2", str(tmp_path), "grid"], env=env) for r in range(2)]
    assert all(p.wait(timeout=160) == 0 for p in procs)
    got = np.load(tmp_path / "gathered.npy")
    drv, n_total = _grid_driver(0, 1)
    assert n_total == 112 and len(drv.chains) == 112
    ref = drv.run(total_sweeps=2, sweep_size=3, start_temp=0.257, perform_annealing=False)
    assert np.array_equal(got[0], ref["energy_hist"]) and np.array_equal(got[2], ref["adsorption_count_hist"])
    # different grid points really see different grand potentials for the same structure
    import bench
    fns = bench.surface_energy_fns("sto_pourbaix", None, [(0.0, -1.0, 0), (14.0, 2.0, 0)])
    sym = ["Sr", "Ti", "O", "O", "O", "H"]
    assert abs(fns[0](-10.0, sym) - fns[1](-10.0, sym)) > 0.1


def test_pipelined_groups_make_the_same_decisions():
    """MultiChainMC.pipeline (host logic of one chain group overlapped with the relaxation of the other,
    asynchronous relax handles) leaves every chain's accept/reject sequence and occupancy unchanged."""
    seeds = list(range(7))
    ref, _ = _driver(seeds)
    for _ in range(12):
        ref.step()
    drv, _ = _driver(seeds)
    sync_fn = drv.relax_fn

    class Handle:          # asynchronous engine call: the result is only read back in step_end
        def __init__(self, out):
            self._out, self.read = out, 0

        def result(self):
            self.read += 1
            return self._out

    handles = []

    def async_fn(pos_l, num_l, fix_l):
        handles.append(Handle(sync_fn(pos_l, num_l, fix_l)))
        return handles[-1]

    drv.relax_fn = async_fn
    pipe = drv.pipeline(n_groups=3)
    flags = [pipe.advance(last=(k == 11)) for k in range(12)]
    pipe.drain()
    assert all(h.read == 1 for h in handles)
    assert drv.decisions == ref.decisions
    for a, b in zip(drv.chains, ref.chains):
        assert np.array_equal(a.occ, b.occ) and np.array_equal(a.arrays()[0], b.arrays()[0])
    assert [f[3] for f in flags] == [d[0] for d in ref.decisions[3]]


def test_checkpoint_resume_and_stats_csv(tmp_path):
    """8f-3: state_dict/load_state_dict (incl. both RNG streams) resumes a run bit-identically at a sweep
    boundary; stats.csv is written in the reference's format."""
    import pickle
    seeds = [3, 4, 5]
    ref, _ = _driver(seeds)
    full = ref.run(total_sweeps=4, sweep_size=5, start_temp=1.0, alpha=0.9, history=True)
    a, _ = _driver(seeds)
    part = a.run(total_sweeps=2, sweep_size=5, start_temp=1.0, alpha=0.9)
    blob = pickle.dumps(a.state_dict())
    b, _ = _driver([99, 98, 97])                      # different seeds: everything must come from the checkpoint
    b.load_state_dict(pickle.loads(blob))
    rest = b.run(total_sweeps=4, sweep_size=5, start_temp=1.0, alpha=0.9, starting_iteration=2)
    assert np.array_equal(part["energy_hist"][:, :2], full["energy_hist"][:, :2])
    assert np.array_equal(rest["energy_hist"][:, 2:], full["energy_hist"][:, 2:])
    assert np.array_equal(rest["adsorption_count_hist"][:, 2:], full["adsorption_count_hist"][:, 2:])
    assert b.decisions == ref.decisions
    assert len(full["history"]) == 4 and np.array_equal(full["history"][-1][1]["occ"], ref.chains[1].occ)
    mc.write_stats_csv(tmp_path / "stats.csv", full, chain=1)
    lines = (tmp_path / "stats.csv").read_text().splitlines()
    assert lines[0] == "energy,frac_accept,adsorption_count" and len(lines) == 5
    assert lines[1] == "%.3f,%.3f,%d" % (full["energy_hist"][1, 0], full["frac_accept_hist"][1, 0], full["adsorption_count_hist"][1, 0])


def test_energy_memo_changes_nothing_but_the_number_of_relaxations():
    """8f-2: occupancy-keyed energy memo / duplicate-proposal coalescing (opt-in)."""
    seeds = list(range(9))
    ref, _ = _driver(seeds)
    for _ in range(15):
        ref.step()
    drv, _ = _driver(seeds)
    drv.energy_memo, drv.memo_hits = {}, 0
    for _ in range(15):
        drv.step()
    assert drv.decisions == ref.decisions
    for a, b in zip(drv.chains, ref.chains):
        assert np.array_equal(a.occ, b.occ)
    assert drv.memo_hits > 0 and drv.n_relaxed + drv.memo_hits == ref.n_relaxed


def test_per_sweep_structure_dumps(tmp_path):
    """SurfaceSystem.save_structures (mcmc/system.py:488-534) -- reference file names; the CIFs read back
    (with the fixture generator's CIF reader) to the dumped coordinates at 5 decimals of the fractional coordinates."""
    import sys
    sys.path.insert(0, str(ROOT / "tests" / "golden"))
    from make_fixtures import load_cif
    drv, (sym0, pos0, sites) = _driver([4])
    drv.cell, drv.pbc = np.diag([6.0, 6.0, 12.0]), True

    def detail(pos_l, num_l, fix_l):
        out = np.zeros((len(pos_l), 8))
        return out, [p + 0.01 for p in pos_l]

    res = drv.run(total_sweeps=2, sweep_size=3, start_temp=0.5, save_folder=tmp_path, save_chains=(0,), relax_detail_fn=detail)
    files = sorted(p.name for p in tmp_path.iterdir())
    assert len(files) == 4 and all(f.startswith("inb_") and f.endswith(".cif") for f in files)
    e1 = res["energy_hist"][0, 0]
    un = [f for f in files if f.startswith("inb_unrelaxed_slab_sweep_001_energy_%.3f_" % e1)]
    assert len(un) == 1
    back = load_cif(tmp_path / files[-1])                       # unrelaxed, sweep 2 = the final MC state
    p, z = drv.chains[0].arrays()
    assert np.array_equal(back["numbers"], z) and np.abs(back["positions"] - p).max() < 12.0 * 1e-5
    rel = load_cif(tmp_path / [f for f in files if "relaxed_slab_sweep_002" in f and "unrelaxed" not in f][0])
    assert np.abs(rel["positions"] - (p + 0.01)).max() < 12.0 * 1e-5
    from surface_sampling_b200 import io
    out = io.write_traj(tmp_path / "t.traj", [(z, p, -1.0), (z, p + 0.1, -1.5)], drv.cell)
    txt = open(out).read().splitlines()
    assert txt[0] == str(len(z)) and "energy=-1.00000000" in txt[1] and len(txt) == 2 * (len(z) + 2)
