"""ORACLE (test infrastructure only — never imported by the product path).

Single-chain restatement of the reference's Monte Carlo loop around the hot path, with the
reference's own RNG call order (global ``np.random`` + stdlib ``random`` seeded once):
  MCMC.run / sweep / step_semigrand / step_canonical / prepare_canonical   mcmc/mcmc.py:150-390
  ChangeProposal / SwitchProposal                                       mcmc/events/proposal.py:74-187
  Change / Exchange forward + acceptance                                 mcmc/events/event.py:52-155
  MetropolisCriterion                                                    mcmc/events/criterion.py:134-168
  change_site / add_atom / add_atom_group / remove_atom                  mcmc/slab.py:235-395
Atoms are kept as a plain Python list of records so ``del`` has ASE's index-shifting semantics.
Pinned by the reference's unit-test known answers (tests/test_mc_state.py ports
tests/test_slab.py:41-74 and tests/test_slab_groups.py:41-87).
"""
from __future__ import annotations

import itertools
import math
import random

import numpy as np

GROUPS = {
    "HO": [("O", (0.0, 0.0, 0.0)), ("H", (1.0, 0.0, 0.0))],
    "H2O": [("O", (0.0, 0.0, 0.0)), ("H", (0.5, -math.sqrt(3) / 2, 0.0)), ("H", (0.5, math.sqrt(3) / 2, 0.0))],
}


def formula(symbols):
    from collections import Counter
    c = Counter(symbols)
    return "".join(f"{k}{c[k] if c[k] > 1 else ''}" for k in sorted(c))  # no carbon on this path


class OracleSurface:
    def __init__(self, symbols, positions, ads_coords, occ=None, ads_group=None):
        grp = list(ads_group) if ads_group is not None else [0] * len(symbols)
        self.atoms = [{"sym": s, "pos": np.array(p, float), "grp": int(g)} for s, p, g in zip(symbols, positions, grp)]
        self.ads_coords = [np.array(c, float) for c in ads_coords]
        self.occ = [0] * len(ads_coords) if occ is None else [int(o) for o in occ]
        self.results = {}

    def copy_state(self):
        return ([dict(a) for a in self.atoms], list(self.occ), dict(self.results))

    def set_state(self, st):
        self.atoms, self.occ, self.results = [dict(a) for a in st[0]], list(st[1]), st[2]

    def start_ads(self, site):
        idx = self.occ[site]
        return [a["sym"] for a in self.atoms if a["grp"] == idx]

    def change_site(self, site, end_ads):
        if site >= len(self.occ):
            raise IndexError("site index out of range")
        if self.occ[site] != 0:
            self.remove(site, self.start_ads(site))
        if end_ads != "None":
            idx = len(self.atoms)
            self.occ[site] = idx
            members = GROUPS[end_ads] if end_ads in GROUPS else [(end_ads, (0.0, 0.0, 0.0))]
            for sym, off in members:
                self.atoms.append({"sym": sym, "pos": self.ads_coords[site] + np.array(off), "grp": idx})

    def remove(self, site, start_ads):
        idx = self.occ[site]
        n = len(start_ads)
        for _ in range(n):
            del self.atoms[idx]
        self.occ = [o - n if o >= idx else o for o in self.occ]
        for a in self.atoms:
            if a["grp"] >= idx:
                a["grp"] -= n
        self.occ = [max(o, 0) for o in self.occ]
        for a in self.atoms:
            a["grp"] = max(a["grp"], 0)
        self.occ[site] = 0

    @property
    def num_adsorbates(self):
        return sum(1 for o in self.occ if o != 0)


def run_chain(seed, symbols0, positions0, ads_coords, adsorbates, energy_fn, total_sweeps, sweep_size,
              start_temp=1.0, alpha=0.99, perform_annealing=True, canonical=False, num_ads_atoms=0):
    """energy_fn(symbols, positions[N,3]) -> surface energy of the RELAXED proposal.
    Returns dict(decisions=[(accept, curr, prev, u)], occ_history, energy_hist, frac_accept_hist, ads_hist)."""
    np.random.seed(seed)
    random.seed(seed)
    surf = OracleSurface(symbols0, positions0, ads_coords)
    decisions, occ_hist = [], []

    def energy():
        return energy_fn([a["sym"] for a in surf.atoms], np.array([a["pos"] for a in surf.atoms]))

    def metropolis(temp, before):
        # criterion.py:134-168
        after = surf.copy_state()
        surf.set_state(before)
        if "surface_energy" not in surf.results:
            surf.results["surface_energy"] = energy()
            before = surf.copy_state()
        prev = surf.results["surface_energy"]
        surf.set_state(after)
        curr = energy()
        surf.results["surface_energy"] = curr
        with np.errstate(over="ignore"):
            p = np.exp(-float(curr - prev) / temp)
        u = np.random.rand()
        acc = bool(u < p)
        if not acc:
            surf.set_state(before)
        decisions.append((acc, curr, prev, u))
        occ_hist.append(list(surf.occ))
        return acc

    def step_semigrand(temp):
        choices = list(adsorbates) + ["None"]
        site = int(np.random.choice(range(len(surf.occ))))
        if surf.occ[site] != 0:
            choices.remove(formula(surf.start_ads(site)))
        else:
            choices.remove("None")
        end = random.choice(choices)
        before = surf.copy_state()
        surf.change_site(site, end)
        return metropolis(temp, before)

    def step_canonical(temp):
        filled = [k for k, o in enumerate(surf.occ) if o != 0]
        curr = {k: list(g) for k, g in itertools.groupby(filled, key=lambda x: surf.atoms[surf.occ[x]]["sym"])}
        empty = [k for k, o in enumerate(surf.occ) if o == 0]
        if empty:
            curr["None"] = empty
        t1, t2 = random.sample(list(curr.keys()), 2)
        s1, s2 = (random.choices(curr[t], weights=np.ones_like(curr[t]), k=1)[0] for t in (t1, t2))
        before = surf.copy_state()
        surf.change_site(s1, t2)
        surf.change_site(s2, t1)
        return metropolis(temp, before)

    temp = start_temp
    if canonical:
        while surf.num_adsorbates < num_ads_atoms:
            step_semigrand(temp)
    if perform_annealing:
        temps = [start_temp * alpha ** k for k in range(total_sweeps)]
        t, temps = start_temp, [start_temp]
        while len(temps) < total_sweeps:
            t *= alpha
            temps.append(t)
    else:
        temps = [start_temp] * total_sweeps
    e_hist, f_hist, a_hist = [], [], []
    for i in range(total_sweeps):
        n_acc = 0
        for _ in range(sweep_size):
            n_acc += step_canonical(temps[i]) if canonical else step_semigrand(temps[i])
        e_hist.append(surf.results["surface_energy"])
        f_hist.append(n_acc / sweep_size)
        a_hist.append(surf.num_adsorbates)
    return {"decisions": decisions, "occ_history": occ_hist, "energy_hist": e_hist, "frac_accept_hist": f_hist,
            "ads_hist": a_hist, "final": surf}
