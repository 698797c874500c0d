"""ORACLE (test infrastructure only — never imported by the product path).

Scalar restatement of the Pourbaix grand potential of ``NFFPourbaix``
(mcmc/calculators/calculators.py:197-305):

  Omega = -(dG1 + dG2)
  dG1 = sum_el n_el * E_std(el) - (E_slab + sum_ads floor(formula / ads) * corr_ads)          :236-274
        (for an adsorbate containing both O and H, max(n_H - n_O, 0) water units are removed
         from the formula first, :254-271)
  dG2 = sum_atoms [ dG2_std(el) - n_e(el) phi - ln(10) n_H(el) kT pH + kT ln c(el) ]          :197-231

PARITY UNPINNED beyond the formula: the reference tree holds no numeric golden for this scalar
(SURVEY.md 8c); the PourbaixAtom table values used in the tests are the literals of
tests/pourbaix/test_pourbaix_atoms.py:44-86 (Sr/O rows) plus synthetic Ti/H rows.
"""
from __future__ import annotations

import math
from collections import Counter


def pourbaix_potential(symbols, slab_energy, table, phi, pH, kT, adsorbate_corrections=None):
    """table[el] = dict(E_std, dG2_std, n_e, n_H, conc)."""
    cnt = Counter(symbols)
    sum_mu = sum(n * table[el]["E_std"] for el, n in cnt.items())
    e = slab_energy
    formula = dict(cnt)
    for ads, corr in (adsorbate_corrections or {}).items():
        ads_cnt = Counter(_parse(ads))
        if "O" in ads_cnt and "H" in ads_cnt:
            extra = max(formula.get("H", 0) - formula.get("O", 0), 0)
            if extra > 0:
                formula = {k: v - {"H": 2 * extra, "O": extra}.get(k, 0) for k, v in formula.items()}
        times = min(formula.get(k, 0) // v for k, v in ads_cnt.items())
        e += max(times, 0) * corr
    dg1 = sum_mu - e
    dg2 = 0.0
    for el in symbols:
        t = table[el]
        dg2 += t["dG2_std"] - t["n_e"] * phi - math.log(10) * t["n_H"] * kT * pH + kT * math.log(t["conc"])
    return -(dg1 + dg2)


def _parse(s):
    import re
    out = []
    for sym, c in re.findall(r"([A-Z][a-z]?)(\d*)", s):
        out += [sym] * (int(c) if c else 1)
    return out
