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


# ---- filter_distances / DistanceCriterion: the reference's tests/test_filter_distance.py:40-97, same vectors ----
def _sto_2x2x1(structures):
    u = structures["SrTiO3_unit_cell"]
    cell = u["cell"].copy()
    pos, num = [], []
    for i in range(2):
        for j in range(2):
            pos.append(u["positions"] + i * cell[0] + j * cell[1])
            num.append(u["numbers"])
    cell[0] *= 2; cell[1] *= 2
    return np.concatenate(pos), np.concatenate(num), cell, u["pbc"]


@pytest.mark.parametrize("ads_pos,expected", [
    ([[1.96777, 1.99250, 18.59954]], False),                                   # one O at a bridge site: too close
    ([[5.90331, 0.14832, 19.49200]], True),                                    # one O on top
    ([[1.96777, 1.99250, 18.59954], [5.90331, 0.14832, 19.49200]], False),     # bridge + top
    ([[5.90331, 0.14832, 19.49200], [1.96777, 4.13332, 19.49200]], True),      # two tops
])
def test_filter_distances_reference_vectors(structures, ads_pos, expected):
    from surface_sampling_b200.mc import filter_distances
    pos, num, cell, pbc = _sto_2x2x1(structures)
    sym = [SYMBOLS[int(z)] for z in num] + ["O"] * len(ads_pos)
    assert filter_distances(sym, np.vstack([pos, ads_pos]), cell, pbc, ads=["O"], cutoff_distance=1.5) is expected


def test_filter_distances_across_the_cell_boundary(structures):
    """tests/test_filter_distance.py:88-97: two O atoms closer than 1.5 A only through the periodic image."""
    from surface_sampling_b200.mc import filter_distances
    s = structures["SrTiO3_001_distance_failed"]
    sym = [SYMBOLS[int(z)] for z in s["numbers"]]
    assert not filter_distances(sym, s["positions"], s["cell"], s["pbc"], ads=["O"], cutoff_distance=1.5)
    assert filter_distances(sym, s["positions"], s["cell"], [False] * 3, ads=["Sr"], cutoff_distance=1.5)


def test_distance_criterion_replaces_metropolis():
    """mcmc/mcmc.py:253-254: with filter_distance > 0 a move is accepted iff the filtered species keep their distance;
    nothing is relaxed and no uniform is drawn."""
    from surface_sampling_b200.mc import MultiChainMC
    calls = []

    def relax_fn(p, z, f):
        calls.append(len(p))
        return np.zeros((len(p), 8))

    sites = np.array([[0.0, 0, 2], [1.0, 0, 2], [4.0, 4, 2]])
    drv = MultiChainMC([NUMBERS["Ga"]] * 2, [[0, 0, 0], [4, 4, 0]], [True, True], sites, ["Sr"], relax_fn, lambda e, s: e, [0, 1],
                       filter_distance=1.5, filter_adsorbate_types=("Sr",), cell=np.eye(3) * 10, pbc=[True, True, False])
    for _ in range(12):
        acc = drv.step()
        for c, a in zip(drv.chains, acc):
            occ = np.flatnonzero(c.occ)
            assert not (0 in occ and 1 in occ)            # sites 0 and 1 are 1.0 A apart: never both filled with Sr
    assert calls == []                                    # no relaxation inside the criterion
    assert any(d[0] for ch in drv.decisions for d in ch) and any(not d[0] for ch in drv.decisions for d in ch)
    drv._ensure_prev(drv.chains)                          # sweep-end energy: evaluated on demand
    assert calls == [2]
