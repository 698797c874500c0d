"""Host-side MC state: known answers ported from the reference's own unit tests
(tests/test_slab.py:41-74, tests/test_slab_groups.py:41-87, tests/events/test_criterion.py) for BOTH the
product state (surface_sampling_b200.mc.ChainState) and the oracle restatement (oracle.mc)."""
import numpy as np
import pytest

from oracle.mc import OracleSurface
from surface_sampling_b200.engine import NUMBERS, SYMBOLS
from surface_sampling_b200.mc import ChainState, create_anneal_schedule, hill_formula, make_site_grid

SYMS = ["Ga", "As", "Ga", "As"]
POS = [[0, 0, 0], [0, 0, 3], [1, 1, 1], [1, 1, 4]]
SITES = [(0, 0, 3), (1, 1, 1), (2, 2, 5)]


def product():
    return ChainState([NUMBERS[s] for s in SYMS], POS, SITES, occ=[1, 2, 0], ads_group0=[0, 1, 2, 0])


def oracle():
    return OracleSurface(SYMS, POS, SITES, occ=[1, 2, 0], ads_group=[0, 1, 2, 0])


def view(s):
    if isinstance(s, ChainState):
        return [SYMBOLS[z] for z in s.numbers], list(s.occ), list(s.ads_group)
    return [a["sym"] for a in s.atoms], list(s.occ), [a["grp"] for a in s.atoms]


@pytest.mark.parametrize("make", [product, oracle])
def test_change_site_known_answers(make):
    s = make(); s.change_site(0, "O")                       # test_slab.py:41-50
    sym, occ, grp = view(s)
    assert len(sym) == 4 and occ == [3, 1, 0] and sym[3] == "O" and grp == [0, 1, 0, 3]
    s = make(); s.change_site(2, "Ir")                      # test_slab.py:53-62
    sym, occ, grp = view(s)
    assert len(sym) == 5 and occ == [1, 2, 4] and sym[4] == "Ir" and grp == [0, 1, 2, 0, 4]
    s = make(); s.change_site(0, "None")                    # test_slab.py:65-74
    sym, occ, grp = view(s)
    assert len(sym) == 3 and occ == [0, 1, 0] and sym[2] == "As" and grp == [0, 1, 0]
    with pytest.raises(IndexError):                         # test_slab.py:77-81
        make().change_site(5, "As")


@pytest.mark.parametrize("make", [product, oracle])
def test_group_sequence_known_answers(make):
    """tests/test_slab_groups.py:41-87 chains state across four calls (module-scoped fixture)."""
    s = make()
    s.change_site(0, "HO")
    sym, occ, grp = view(s)
    assert len(sym) == 5 and occ == [3, 1, 0] and sym[3:] == ["O", "H"] and grp == [0, 1, 0, 3, 3]
    s.change_site(2, "Ir")
    sym, occ, grp = view(s)
    assert len(sym) == 6 and occ == [3, 1, 5] and sym[5] == "Ir" and grp == [0, 1, 0, 3, 3, 5]
    s.change_site(0, "None")
    sym, occ, grp = view(s)
    assert len(sym) == 4 and occ == [0, 1, 3] and sym[3] == "Ir" and grp == [0, 1, 0, 3]
    s.change_site(1, "None")
    sym, occ, grp = view(s)
    assert len(sym) == 3 and occ == [0, 0, 2] and sym[2] == "Ir" and grp == [0, 0, 2]


def test_group_geometry_and_formula():
    s = product()
    s.change_site(2, "H2O")
    p = np.array(s.positions[-3:])
    assert np.allclose(p[0], SITES[2]) and np.allclose(np.linalg.norm(p[1:] - p[0], axis=1), 1.0)
    assert hill_formula(s.symbols_at_site(2)) == "H2O" and hill_formula(["O", "H"]) == "HO"


def test_proposals_replay_reference_rng_order():
    """Chain c must consume exactly the draws the single-chain reference consumes after
    np.random.seed(s); random.seed(s)  (SURVEY.md App. A.6)."""
    import random
    seed = 11
    c = ChainState([NUMBERS[s] for s in SYMS], POS, SITES, seed=seed)
    np.random.seed(seed); random.seed(seed)
    for _ in range(20):
        a = c.propose_change(["Sr", "O"])
        site = np.random.choice(range(3))
        choices = ["Sr", "O", "None"]
        choices.remove("None" if c.occ[site] == 0 else hill_formula(c.symbols_at_site(site)))
        end = random.choice(choices)
        assert (a["site_idx"], a["end_ads"]) == (site, end)
        c.apply(a)
        assert c.np_rng.rand() == np.random.rand()


def test_anneal_schedule_and_sites(structures):
    t = create_anneal_schedule(1.0, 5, 0.99)
    assert np.allclose(t, [1.0, 0.99, 0.99 ** 2, 0.99 ** 3, 0.99 ** 4])
    s = structures["SrTiO3_001_2x2"]
    g = make_site_grid(s["positions"], s["cell"], 64, 1.5)
    assert g.shape == (64, 3) and np.allclose(g[:, 2], s["positions"][:, 2].max() + 1.5)
    assert len({tuple(np.round(x, 6)) for x in g}) == 64
