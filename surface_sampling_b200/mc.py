"""Multi-chain VSSR-MC driver: the callers on either side of the hot path, batched over chains.

Reference semantics kept exactly:
  * state per chain = ``occ`` (atom index adsorbed at each virtual site, 0 = empty), appended
    adsorbate atoms with ASE index semantics (append at end; deletion shifts higher indices) and
    the ``ads_group`` tag — mcmc/slab.py:235-395, mcmc/system.py:149-182;
  * RNG draw order per iteration — semigrand: ``np.random.choice(range(n_sites))`` ->
    ``random.choice(ads_choices)`` -> ``np.random.rand()`` (events/proposal.py:85,99,
    events/criterion.py:168); canonical: ``random.sample(types, 2)`` -> ``random.choices`` x2 ->
    ``np.random.rand()`` (slab.py:70,219-222).  Every chain owns a private legacy
    ``np.random.RandomState(seed)`` and ``random.Random(seed)``, so chain c replays exactly what the
    single-chain reference does after ``np.random.seed(seed); random.seed(seed)``;
  * every proposal is relaxed from the IDEAL-site structure (system.py:348-357,370); Metropolis
    compares SURFACE energies; one uniform is drawn per iteration even when p >= 1; the very
    first iteration recomputes the missing ``prev`` energy (criterion.py:144-149).
All chains advance in lock step: one batched relax call per MC iteration, only 8 scalars per chain
come back from the GPU.
"""
from __future__ import annotations

import itertools
import math
import random
from collections import Counter

import numpy as np

from .engine import NUMBERS, SYMBOLS

# mcmc/slab.py:22-32
ATOM_GROUPS = {
    "HO": (["O", "H"], np.array([[0.0, 0.0, 0.0], [1.0, 0.0, 0.0]])),
    "H2O": (["O", "H", "H"], np.array([[0.0, 0.0, 0.0], [0.5, -math.sqrt(3) / 2, 0.0], [0.5, math.sqrt(3) / 2, 0.0]])),
}


def hill_formula(symbols) -> str:
    c = Counter(symbols)
    keys = sorted(c)
    if "C" in c:
        keys = ["C"] + (["H"] if "H" in c else []) + [k for k in keys if k not in ("C", "H")]
    return "".join(f"{k}{c[k] if c[k] > 1 else ''}" for k in keys)


_SYMBOLS_ARR = np.array([None] + [SYMBOLS[z] for z in range(1, max(SYMBOLS) + 1)], dtype=object)


class ChainState:
    """SurfaceSystem state restricted to what the MC loop mutates."""

    def __init__(self, numbers0, positions0, ads_coords, occ=None, seed=0, ads_group0=None):
        self.n0 = len(numbers0)
        self.numbers = list(int(z) for z in numbers0)          # real_atoms numbers
        self.positions = [np.asarray(p, dtype=float) for p in positions0]
        # the MC state keeps the UNRELAXED structure (relaxed coordinates are never written back), so the
        # first n0 rows never change: arrays() only converts the adsorbate tail
        self._pos0 = np.array(self.positions, dtype=np.float64).reshape(-1, 3)
        self._num0 = np.array(self.numbers, dtype=np.int64)
        self.ads_group = [0] * self.n0 if ads_group0 is None else [int(g) for g in ads_group0]
        self.ads_coords = np.asarray(ads_coords, dtype=float)
        self.occ = np.zeros(len(self.ads_coords), dtype=int) if occ is None else np.array(occ, dtype=int)
        self.results = {}
        self.np_rng = np.random.RandomState(seed)
        self.py_rng = random.Random(seed)

    def __len__(self):
        return len(self.numbers)

    # ---- save/restore (system.py:149-182) ----
    def snapshot(self):
        return (list(self.numbers), list(self.positions), list(self.ads_group), self.occ.copy(), dict(self.results))

    def restore(self, snap, adopt=False):
        """adopt=True takes over the snapshot's containers instead of copying them (the caller drops the snapshot)."""
        if adopt:
            self.numbers, self.positions, self.ads_group, self.occ, self.results = snap
            return
        self.numbers, self.positions, self.ads_group, self.occ, self.results = (
            list(snap[0]), list(snap[1]), list(snap[2]), snap[3].copy(), dict(snap[4]))

    # ---- checkpoint (SurfaceSystem.todict/fromdict, system.py:591-653; here incl. both RNG states, which the
    # reference does not checkpoint: a resumed chain continues bit-identically) ----
    def todict(self) -> dict:
        return {"n0": self.n0, "numbers": list(self.numbers), "positions": np.array(self.positions, dtype=float).reshape(-1, 3),
                "ads_group": list(self.ads_group), "ads_coords": self.ads_coords.copy(), "occ": self.occ.copy(),
                "results": dict(self.results), "np_rng": self.np_rng.get_state(), "py_rng": self.py_rng.getstate()}

    @classmethod
    def fromdict(cls, d: dict) -> "ChainState":
        n0 = int(d["n0"])
        c = cls(d["numbers"][:n0], d["positions"][:n0], d["ads_coords"], occ=d["occ"], ads_group0=d["ads_group"][:n0])
        c.numbers = [int(z) for z in d["numbers"]]
        c.positions = [np.asarray(p, dtype=float) for p in d["positions"]]
        c.ads_group = [int(g) for g in d["ads_group"]]
        c.results = dict(d["results"])
        c.np_rng.set_state(d["np_rng"])
        c.py_rng.setstate(d["py_rng"])
        return c

    @property
    def num_adsorbates(self):
        return int(np.count_nonzero(self.occ))

    def symbols_at_site(self, site_idx):
        """get_start_ads (slab.py:277-289): all atoms whose ads_group == occ[site]."""
        idx = int(self.occ[site_idx])
        if idx == 0:      # empty site: the reference's comparison then matches the slab atoms (group 0)
            return [SYMBOLS[self.numbers[a]] for a in range(len(self)) if self.ads_group[a] == 0]
        # an adsorbate group is a contiguous block starting at the atom whose index is the group id
        ag, out, a = self.ads_group, [], idx
        while a < len(ag) and ag[a] == idx:
            out.append(SYMBOLS[self.numbers[a]])
            a += 1
        return out

    # ---- slab.py:235-395 ----
    def change_site(self, site_idx, end_ads):
        if site_idx >= len(self.occ):
            raise IndexError("site index out of range")
        if self.occ[site_idx] != 0:
            self.remove_atom(site_idx, self.symbols_at_site(site_idx))
        if end_ads != "None":
            if end_ads in ATOM_GROUPS:
                self.add_atom_group(site_idx, end_ads)
            else:
                self.add_atom(site_idx, end_ads)

    def add_atom(self, site_idx, adsorbate):
        idx = len(self)
        self.occ[site_idx] = idx
        self.numbers.append(NUMBERS[adsorbate])
        self.positions.append(self.ads_coords[site_idx].copy())
        self.ads_group.append(idx)

    def add_atom_group(self, site_idx, group_name):
        if group_name not in ATOM_GROUPS:
            raise ValueError(f"Unknown group name: {group_name}")
        syms, offs = ATOM_GROUPS[group_name]
        idx = len(self)
        self.occ[site_idx] = idx
        for s, o in zip(syms, offs):
            self.numbers.append(NUMBERS[s])
            self.positions.append(self.ads_coords[site_idx] + o)
            self.ads_group.append(idx)

    def remove_atom(self, site_idx, start_ads):
        idx = int(self.occ[site_idx])
        hits = int((self.occ == idx).sum())
        assert hits == 1, "no sites found" if hits == 0 else "more than 1 site found"
        n = len(start_ads)
        if n > 1 and hill_formula(start_ads) not in ATOM_GROUPS:
            raise ValueError(f"Unknown group name: {hill_formula(start_ads)}")
        del self.numbers[idx:idx + n]
        del self.positions[idx:idx + n]
        del self.ads_group[idx:idx + n]
        # slab.py:372-395 shifts every index >= idx down by n and clamps negatives to 0.  A group id is the index of
        # its first atom, so only entries behind the removed block can be >= idx, and idx >= n0 > n keeps them positive.
        occ = self.occ
        occ = np.where(occ >= idx, occ - n, occ)      # (a fresh array: snapshots keep theirs)
        if n >= idx:                                  # never with idx >= n0 > n; kept for the reference's clamp
            occ[occ < 0] = 0
        occ[site_idx] = 0
        self.occ = occ
        ag = self.ads_group
        for k in range(idx, len(ag)):
            g = ag[k]
            if g >= idx:
                ag[k] = g - n if g - n > 0 else 0

    # ---- proposals ----
    def propose_change(self, adsorbates, site_idx=None):
        """ChangeProposal.get_action (events/proposal.py:74-106)."""
        choices = list(adsorbates) + ["None"]
        if site_idx is None:
            site_idx = int(self.np_rng.randint(0, len(self.occ)))   # == choice(range(n)): same draw, 10x cheaper
        if self.occ[site_idx] != 0:
            start = hill_formula(self.symbols_at_site(site_idx))
            choices.remove(start)
        else:
            start = "None"
            choices.remove("None")
        end = self.py_rng.choice(choices)
        return {"name": "change", "site_idx": site_idx, "start_ads": start, "end_ads": end}

    def propose_switch(self):
        """SwitchProposal.get_action -> get_complementary_idx with uniform weights (slab.py:168-232).
        Note the reference's groupby-into-dict keeps only the LAST run of each symbol; replicated."""
        numbers = self.numbers
        # {symbol: sites of its LAST run} in first-appearance order of the symbols -- what the reference's
        # {k: list(g) for k, g in groupby(filled, key=symbol)} builds (a repeated key keeps its slot, takes the new
        # run) -- in one pass over the sites
        curr, empty, last, run = {}, [], None, None
        for x, o in enumerate(self.occ.tolist()):
            if o == 0:
                empty.append(x)
                continue
            sym = SYMBOLS[numbers[o]]
            if sym == last:
                run.append(x)
            else:
                run = [x]
                curr[sym] = run
                last = sym
        if empty:
            curr["None"] = empty
        t1, t2 = self.py_rng.sample(list(curr.keys()), 2)
        # random.choices(pop, weights=ones, k=1) draws ONE random() and bisects the cumulative weights 1..n at
        # r*n, i.e. picks pop[floor(r*n)]: same draw, same element, without building numpy weight arrays
        s1, s2 = (curr[t][int(self.py_rng.random() * len(curr[t]))] for t in (t1, t2))
        return {"name": "switch", "site1_idx": s1, "site2_idx": s2, "site1_ads": t1, "site2_ads": t2}

    def apply(self, action):
        if action["name"] == "change":
            self.change_site(action["site_idx"], action["end_ads"])
        else:  # Exchange.forward (events/event.py:138-155)
            self.change_site(action["site1_idx"], action["site2_ads"])
            self.change_site(action["site2_idx"], action["site1_ads"])

    def arrays(self):
        n0 = self.n0
        if len(self.numbers) == n0:
            return self._pos0.copy(), self._num0.copy()
        tail = np.array(self.positions[n0:], dtype=np.float64).reshape(-1, 3)
        return np.concatenate([self._pos0, tail]), np.concatenate([self._num0, np.array(self.numbers[n0:], dtype=np.int64)])


def all_distances_mic(positions, cell, pbc) -> np.ndarray:
    """ase ``Atoms.get_all_distances(mic=True)``: minimum-image distance matrix (general cell: the 27 neighbouring
    images of the wrapped difference are searched along the periodic directions)."""
    pos = np.asarray(positions, dtype=float)
    cell = np.asarray(cell, dtype=float).reshape(3, 3)
    pbc = np.broadcast_to(np.asarray(pbc, dtype=bool), (3,))
    d = pos[None, :, :] - pos[:, None, :]
    if not pbc.any() or abs(np.linalg.det(cell)) < 1e-12:
        return np.linalg.norm(d, axis=-1)
    frac = d @ np.linalg.inv(cell)
    frac[..., pbc] -= np.round(frac[..., pbc])
    d = frac @ cell
    rng = [(-1, 0, 1) if p else (0,) for p in pbc]
    best = np.full(d.shape[:2], np.inf)
    for i in rng[0]:
        for j in rng[1]:
            for k in rng[2]:
                best = np.minimum(best, np.linalg.norm(d + np.array([i, j, k], float) @ cell, axis=-1))
    return best


def filter_distances(symbols, positions, cell, pbc, ads=("O",), cutoff_distance: float = 1.5) -> bool:
    """mcmc/utils/misc.py:118-135: True iff no two atoms of the `ads` species are closer than `cutoff_distance`
    (minimum image).  Pinned by the reference's tests/test_filter_distance.py:40-97 (tests/test_mc_state.py)."""
    sel = np.isin(np.asarray(symbols, dtype=object), list(ads))
    if sel.sum() < 2:
        return True
    dist = np.triu(all_distances_mic(np.asarray(positions, float)[sel], cell, pbc))
    return not np.any((dist > 0) & (dist <= cutoff_distance))


def create_anneal_schedule(start_temp=1.0, total_sweeps=1000, alpha=0.99):
    """mcmc/utils/sampling.py:10-71 (single-anneal branch)."""
    temps = [start_temp]
    t = start_temp
    while len(temps) < total_sweeps:
        t *= alpha
        temps.append(t)
    return temps[:total_sweeps]


def make_site_grid(positions, cell, n_sites, height=1.5):
    """Deterministic synthetic virtual-site grid (pymatgen's AdsorbateSiteFinder is unavailable:
    SURVEY.md 8d): nx x ny fractional grid at z_top + height, truncated to n_sites."""
    cell = np.asarray(cell, dtype=float)
    ratio = np.linalg.norm(cell[0]) / np.linalg.norm(cell[1])
    nx = max(1, int(round(math.sqrt(n_sites * ratio))))
    ny = int(math.ceil(n_sites / nx))
    z = float(np.asarray(positions)[:, 2].max()) + height
    pts = []
    for a in range(nx):
        for b in range(ny):
            f = np.array([(a + 0.25) / nx, (b + 0.25) / ny, 0.0])
            p = f @ cell
            p[2] = z
            pts.append(p)
    return np.array(pts[:n_sites])


class MultiChainMC:
    """C independent chains in lock step over one batched engine.

    ``relax_fn(pos_list, num_list, fixed_list) -> out[C,8]`` is the hot-path call (engine relax);
    ``surface_energy_fn(energy, symbols) -> float`` is the host scalar (H6/H7) -- one callable for every chain, or a
    list with one callable per chain (Pourbaix runs: every chain carries its own (pH, U) grid point, BASELINE
    config 5 / scripts/sample_pourbaix_surface.py:253-259).  ``occ0`` / ``ads_group0`` start the chains from sampled
    surface atoms (``sample_surface_atoms``, scripts/sample_pourbaix_surface.py:218-236)."""

    def __init__(self, numbers0, positions0, fixed0, ads_coords, adsorbates, relax_fn, surface_energy_fn, seeds,
                 canonical=False, num_ads_atoms=0, occ0=None, energy_memo=False, ads_group0=None, filter_distance=0.0,
                 filter_adsorbate_types=("Sr", "Ti"), cell=None, pbc=True):
        # filter_distance > 0 (mcmc/mcmc.py:218-219,253-254): the DistanceCriterion REPLACES Metropolis -- a move is
        # accepted iff no two atoms of `filter_adsorbate_types` (criterion.py:93-97 default) end up closer than that
        # distance; no relaxation, no uniform draw; the sweep-end energy is evaluated as usual
        self.filter_distance, self.filter_adsorbate_types = float(filter_distance), tuple(filter_adsorbate_types)
        self.cell, self.pbc = cell, pbc
        if self.filter_distance > 0:
            assert cell is not None, "filter_distance needs the cell (minimum-image distances)"
        # energy_memo (off by default and never used by bench.py): the relaxed result is a pure
        # function of the unrelaxed structure, and the engine is batch-invariant, so identical structures -- across
        # chains in one step or revisited later -- are relaxed once and their 8 scalars reused bit for bit
        self.energy_memo = {} if energy_memo else None
        self.memo_hits = 0
        self.adsorbates = list(adsorbates)
        self.relax_fn, self.surface_energy_fn = relax_fn, surface_energy_fn
        self.fixed0 = np.asarray(fixed0, dtype=bool)
        self.chains = [ChainState(numbers0, positions0, ads_coords, occ=occ0, seed=s, ads_group0=ads_group0) for s in seeds]
        self._se_per_chain = isinstance(surface_energy_fn, (list, tuple))
        if self._se_per_chain:
            assert len(surface_energy_fn) == len(self.chains), "one surface_energy_fn per chain"
        for k, c in enumerate(self.chains):
            c.index = k
        self.canonical, self.num_ads_atoms = canonical, num_ads_atoms
        if canonical:
            assert num_ads_atoms > 0, "for canonical runs, need number of adsorbed atoms greater than 0"
        self.temp = 1.0
        self.n_relaxed = 0
        self.decisions = [[] for _ in seeds]        # per chain: (accept, curr, prev, u)
        self._fixed_masks = {}

    # -- hot path call ------------------------------------------------------------------------
    # relax_fn may return out[C,8] directly, or a handle with .result() -> out[C,8] when the engine call is
    # asynchronous (the relaxation is enqueued on the GPU and only result() waits for the 8 scalars per chain)
    def _launch(self, chains):
        pos, num, fix = [], [], []
        masks = self._fixed_masks
        for c in chains:
            p, z = c.arrays()
            pos.append(p)
            num.append(z)
            m = masks.get(len(z))
            if m is None:       # the frozen mask depends on the atom count only (adsorbates are appended, never frozen)
                m = masks[len(z)] = np.concatenate([self.fixed0, np.zeros(len(z) - len(self.fixed0), dtype=bool)])
            fix.append(m)
        fns = [self.surface_energy_fn[c.index] for c in chains] if self._se_per_chain else None
        if self.energy_memo is not None:
            return self._launch_memo(pos, num, fix), num, fns
        handle = self.relax_fn(pos, num, fix)
        self.n_relaxed += len(chains)
        return handle, num, fns

    def _launch_memo(self, pos, num, fix):
        keys = [(z.tobytes(), p.tobytes()) for p, z in zip(pos, num)]
        todo, seen = [], {}
        for k, key in enumerate(keys):
            if key not in self.energy_memo and key not in seen:
                seen[key] = len(todo)
                todo.append(k)
        handle = self.relax_fn([pos[k] for k in todo], [num[k] for k in todo], [fix[k] for k in todo]) if todo else None
        self.n_relaxed += len(todo)
        self.memo_hits += len(keys) - len(todo)
        memo = self.energy_memo

        class Joined:
            def result(_self):
                if handle is not None:
                    fresh = handle.result() if hasattr(handle, "result") else handle
                    for key, row in seen.items():
                        memo[key] = np.array(fresh[row], dtype=np.float64)
                return np.stack([memo[key] for key in keys])

        return Joined()

    def _collect(self, launched):
        handle, num, fns = launched
        out = handle.result() if hasattr(handle, "result") else handle
        # the OOB clamp of optimize_slab is invisible to Metropolis (system.py:466-469): raw energy is used
        if fns is None:
            return [self.surface_energy_fn(float(out[k, 2]), _SYMBOLS_ARR[num[k]].tolist()) for k in range(len(num))]
        return [fns[k](float(out[k, 2]), _SYMBOLS_ARR[num[k]].tolist()) for k in range(len(num))]

    def _energies(self, chains):
        return self._collect(self._launch(chains))

    def _ensure_prev(self, chains):
        missing = [c for c in chains if "surface_energy" not in c.results]
        if missing:
            for c, e in zip(missing, self._energies(missing)):
                c.results["surface_energy"] = e

    def step_begin(self, active=None, force_semigrand=None):
        """First half of an MC iteration for the chains in `active` (default all): propose, apply and
        enqueue the relaxation.  Returns a ticket for step_end; nothing here waits for the GPU (unless a
        chain has no current energy yet), so the host work of one chain group overlaps the relaxation of
        another (run_pipelined)."""
        idx = list(range(len(self.chains))) if active is None else list(active)
        chains = [self.chains[i] for i in idx]
        if self.filter_distance > 0:
            return self._step_distance(idx, chains, force_semigrand)
        snaps, actions = [], []
        for k, c in enumerate(chains):
            semigrand = (not self.canonical) or (force_semigrand is not None and force_semigrand[k])
            action = c.propose_change(self.adsorbates) if semigrand else c.propose_switch()
            snaps.append(c.snapshot())
            actions.append(action)
        # criterion: prev from the "before" state (recomputed when missing), curr from "after"
        self._ensure_prev(chains)
        prev = [c.results["surface_energy"] for c in chains]
        for c, a in zip(chains, actions):
            c.apply(a)
        return idx, chains, snaps, prev, self._launch(chains), self.temp

    def _step_distance(self, idx, chains, force_semigrand):
        accepts = []
        for k, c in enumerate(chains):
            semigrand = (not self.canonical) or (force_semigrand is not None and force_semigrand[k])
            action = c.propose_change(self.adsorbates) if semigrand else c.propose_switch()
            snap = c.snapshot()
            c.apply(action)
            p, z = c.arrays()
            acc = filter_distances(_SYMBOLS_ARR[z], p, self.cell, self.pbc, self.filter_adsorbate_types, self.filter_distance)
            if acc:
                c.results.pop("surface_energy", None)      # the state changed without an energy: recomputed when needed
            else:
                c.restore(snap)
            self.decisions[idx[k]].append((acc, None, None, None))
            accepts.append(acc)
        return ("done", accepts)

    def step_end(self, ticket):
        """Second half: read the relaxed energies back and apply Metropolis. Returns accept flags."""
        if ticket[0] == "done":
            return ticket[1]
        idx, chains, snaps, prev, launched, temp = ticket
        curr = self._collect(launched)
        accepts = []
        with np.errstate(over="ignore"):
            probs = [np.exp(-float(curr[k] - prev[k]) / temp) for k in range(len(chains))]
        for k, c in enumerate(chains):
            base_prob = probs[k]
            u = c.np_rng.rand()
            acc = bool(u < base_prob)
            if acc:
                c.results["surface_energy"] = curr[k]
            else:
                snap = snaps[k]
                snap[4]["surface_energy"] = prev[k]
                c.restore(snap, adopt=True)      # the ticket's snapshot is not used again
            self.decisions[idx[k]].append((acc, curr[k], prev[k], u))
            accepts.append(acc)
        return accepts

    def step(self, active=None, force_semigrand=None):
        """One MC iteration for the chains in `active` (default all). Returns accept flags."""
        return self.step_end(self.step_begin(active, force_semigrand))

    def pipeline(self, n_groups=2):
        """Software pipeline over `n_groups` interleaved chain groups: `advance()` completes one MC iteration
        of EVERY chain while the next iteration of each group is already enqueued, so the GPU never waits for
        the host-side Metropolis / proposal logic.  Chains are independent and the engine is batch-invariant,
        so every chain makes exactly the decisions it makes under step()."""
        return _Pipeline(self, n_groups)

    def prepare_canonical(self):
        """MCMC.prepare_canonical (mcmc.py:150-188): semigrand steps until num_ads_atoms adsorbed."""
        while True:
            todo = [i for i, c in enumerate(self.chains) if c.num_adsorbates < self.num_ads_atoms]
            if not todo:
                return
            self.step(active=todo, force_semigrand=[True] * len(todo))

    def state_dict(self) -> dict:
        """Everything needed to resume: chain states with their RNG streams, temperature, counters."""
        return {"chains": [c.todict() for c in self.chains], "temp": self.temp, "n_relaxed": self.n_relaxed,
                "decisions": [list(d) for d in self.decisions]}

    def load_state_dict(self, sd: dict):
        self.chains = [ChainState.fromdict(d) for d in sd["chains"]]
        for k, c in enumerate(self.chains):
            c.index = k
        self.temp, self.n_relaxed = sd["temp"], sd["n_relaxed"]
        self.decisions = [list(d) for d in sd["decisions"]]

    def run(self, total_sweeps=10, sweep_size=20, start_temp=1.0, perform_annealing=True, alpha=0.99,
            anneal_schedule=None, gather=None, starting_iteration=0, history=False, save_folder=None, save_chains=(0,),
            relax_detail_fn=None):
        """MCMC.run (mcmc.py:301-390) for every chain; returns per-chain histories
        (energy_hist, frac_accept_hist, adsorption_count_hist) as [C, total_sweeps] arrays.
        `starting_iteration` resumes a run (mcmc.py:313,381) after load_state_dict; `history=True` also returns
        the per-sweep chain snapshots (`results["history"]`, scripts/sample_surface.py:204-208) as todict()s.
        `save_folder`: after every sweep the chains in `save_chains` are dumped like SurfaceSystem.save_structures
        (mcmc/system.py:488-534, called from mcmc.py:289): unrelaxed CIF, and -- when `relax_detail_fn(pos_list, num_list,
        fix_list) -> (out[C,8], relaxed_positions_list)` is given -- the relaxed CIF (one extra batched relaxation per
        sweep, as the reference spends <= 2 extra evaluations per sweep there)."""
        self.temp = start_temp
        if self.canonical and starting_iteration == 0:
            self.prepare_canonical()
        if anneal_schedule is not None:
            temps = list(anneal_schedule)
        elif perform_annealing:
            temps = create_anneal_schedule(start_temp, total_sweeps, alpha)
        else:
            temps = [start_temp] * total_sweeps
        C = len(self.chains)
        energy_hist = np.zeros((C, total_sweeps))
        frac_accept = np.zeros((C, total_sweeps))
        ads_count = np.zeros((C, total_sweeps), dtype=int)
        snapshots = []
        for i in range(int(starting_iteration), total_sweeps):
            self.temp = temps[i]
            n_acc = np.zeros(C)
            for _ in range(sweep_size):
                n_acc += np.array(self.step(), dtype=float)
            self._ensure_prev(self.chains)
            energy_hist[:, i] = [c.results["surface_energy"] for c in self.chains]
            frac_accept[:, i] = n_acc / sweep_size
            ads_count[:, i] = [c.num_adsorbates for c in self.chains]
            if history:
                snapshots.append([c.todict() for c in self.chains])
            if save_folder is not None:
                self.save_structures(save_folder, i + 1, save_chains, relax_detail_fn)
            if gather is not None:
                gather(i, energy_hist[:, i], frac_accept[:, i], ads_count[:, i])
        out = {"energy_hist": energy_hist, "frac_accept_hist": frac_accept, "adsorption_count_hist": ads_count}
        if history:
            out["history"] = snapshots
        return out


    def save_structures(self, save_folder, sweep_num, chains=(0,), relax_detail_fn=None):
        """Per-sweep dumps of the selected chains (io.save_structures). Returns the written paths."""
        from pathlib import Path

        from . import io
        if self.cell is None:
            raise ValueError("save_structures needs the cell: MultiChainMC(..., cell=..., pbc=...)")
        sel = [self.chains[k] for k in chains]
        arrs = [c.arrays() for c in sel]
        fix = [np.concatenate([self.fixed0, np.zeros(len(z) - len(self.fixed0), dtype=bool)]) for _, z in arrs]
        relaxed, oob = [None] * len(sel), [False] * len(sel)
        if relax_detail_fn is not None:
            out, relaxed = relax_detail_fn([p for p, _ in arrs], [z for _, z in arrs], fix)
            oob = [bool(out[k, 6]) for k in range(len(sel))]
        written = []
        for k, c in enumerate(sel):
            folder = Path(save_folder) if len(self.chains) == 1 else Path(save_folder) / f"chain_{chains[k]:04d}"
            written += io.save_structures(folder, sweep_num, c.results["surface_energy"], arrs[k][1], arrs[k][0], self.cell,
                                          relaxed_pos=relaxed[k], energy_oob=oob[k],
                                          pbc=np.broadcast_to(np.asarray(self.pbc, dtype=bool), (3,)))
        return written


class _Pipeline:
    def __init__(self, drv: MultiChainMC, n_groups: int):
        C = len(drv.chains)
        self.drv = drv
        self.groups = [list(range(g, C, n_groups)) for g in range(n_groups) if g < C]
        self.tickets = [None] * len(self.groups)

    def advance(self, last=False):
        """One iteration of every chain. Returns the accept flags in chain order."""
        drv = self.drv
        acc = [False] * len(drv.chains)
        for gi, g in enumerate(self.groups):
            if self.tickets[gi] is None:
                self.tickets[gi] = drv.step_begin(g)
        for gi, g in enumerate(self.groups):
            flags = drv.step_end(self.tickets[gi])
            self.tickets[gi] = None if last else drv.step_begin(g)
            for i, f in zip(g, flags):
                acc[i] = f
        return acc

    def drain(self):
        """Finish the iterations that are still in flight."""
        for gi in range(len(self.groups)):
            if self.tickets[gi] is not None:
                self.drv.step_end(self.tickets[gi])
                self.tickets[gi] = None


def write_stats_csv(path, results: dict, chain: int = 0):
    """stats.csv of one chain in the reference's format (scripts/sample_surface.py:220-229: columns energy,
    frac_accept, adsorption_count; float_format %.3f, no index)."""
    e, f, a = (np.asarray(results[k])[chain] for k in ("energy_hist", "frac_accept_hist", "adsorption_count_hist"))
    with open(path, "w") as fh:
        fh.write("energy,frac_accept,adsorption_count\n")
        for k in range(len(e)):
            fh.write("%.3f,%.3f,%d\n" % (e[k], f[k], int(a[k])))
