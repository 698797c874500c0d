"""Size-independent properties at BASELINE.json's FULL sizes (where the CPU oracle would take hours).

PaiNN ensemble, 128 chains x ~92 atoms (the burnt-in coverage of the headline workload): Newton's third law, rigid
translation, atom permutation, rotation equivariance, memo/constrained vs plain engine, bitwise equality of duplicated
chains.  Classical potentials at 256 (GaN, config 2) and 1024 chains: translation, force sums, duplicate chains.
Neighbour list of the full batch: symmetric, receiver-sorted, cutoff-complete against brute force on sampled rows.

Tolerances are the north star's (1e-5 eV/atom, 1e-4 eV/A in fp32), written where they are used; transformed inputs
are different fp32 round-off paths of the same arithmetic, so each side may sit anywhere inside its own band."""
import numpy as np
import pytest
import torch

from conftest import perturbed
from oracle import relax as orelax

pytestmark = pytest.mark.gpu
PBC3 = np.array([True, True, True])
E_TOL_PER_ATOM = 1e-5   # eV/atom
F_TOL = 1e-4            # eV/A
C_FULL = 128            # chains per GPU of config 4 (1024 chains over 8 GPUs)


def _late_chains(structures, n_chains, rng, lo=26, hi=40):
    """Slab + `lo..hi` adsorbates of random species on distinct slots of three 4x4 layers (1.9 A apart, 1.5 A above the
    surface -- physical distances, unlike the bench's 1 A site grid), everything displaced by 0.03 A: the atom counts
    and edge counts of the stationary regime of the headline workload."""
    s = structures["SrTiO3_001_2x2"]
    ztop = s["positions"][:, 2].max()
    slots = np.array([[0.5 + 1.9 * (a % 4) + 0.9 * ((a // 16) % 2), 0.5 + 1.9 * ((a // 4) % 4) + 0.9 * ((a // 16) % 2),
                       ztop + 1.5 + 1.9 * (a // 16)] for a in range(48)])
    out = []
    for _ in range(n_chains):
        k = int(rng.integers(lo, hi + 1))
        pick = np.sort(rng.choice(len(slots), size=k, replace=False))
        z = rng.choice([38, 22, 8], size=k)
        pos = np.vstack([s["positions"], slots[pick]])
        out.append(perturbed({"positions": pos, "numbers": np.concatenate([s["numbers"], z]), "cell": s["cell"]}, rng, 0.03))
    return out


def _batch(structs, fixed=None):
    from surface_sampling_b200 import engine
    return engine.Batch.from_arrays([c["positions"] for c in structs], [c["numbers"] for c in structs],
                                    [c["cell"] for c in structs], [PBC3 for _ in structs], fixed)


def _run(eng, structs, fixed=None):
    b = _batch(structs, fixed)
    r = eng.energy_forces(b)
    torch.cuda.synchronize()
    return r["energy"].cpu().numpy().astype(np.float64), b.split_host(r["forces"].cpu().numpy().astype(np.float64)), r, b


@pytest.fixture(scope="module")
def painn_full(structures):
    from surface_sampling_b200 import engine, loaders
    states = [loaders.init_random_weights(q) for q in (0, 1, 2)]
    chains = _late_chains(structures, C_FULL, np.random.default_rng(2026))
    eng = engine.PainnEngine(states, None)
    e, f, _, _ = _run(eng, chains)
    n = np.array([len(c["numbers"]) for c in chains])
    assert n.min() >= 86 and n.max() <= 100 and abs(n.mean() - 93) < 2          # the regime bench.py reports
    # the comparisons below are meaningful only on physical structures (no overlapping placements)
    fmax = np.array([np.abs(x).max() for x in f])
    assert np.median(fmax) < 50, np.median(fmax)
    return {"states": states, "chains": chains, "eng": eng, "e": e, "f": f, "n": n, "fmax": fmax}


def _ftol(fmax):
    return F_TOL + (2e-6 * fmax if fmax > 50 else 0.0)


def test_painn_forces_sum_to_zero_full_size(painn_full):
    """dE/dx from the hand-written backward is a sum of equal and opposite pair terms: the net force on every
    structure vanishes up to fp32 round-off (each of the ~5000 edge terms carries one ulp of the largest force)."""
    for k, f in enumerate(painn_full["f"]):
        net = np.abs(f.sum(axis=0)).max()
        assert net <= 2 * F_TOL + 5e-6 * painn_full["fmax"][k] * np.sqrt(len(f)), (k, net, painn_full["fmax"][k])


def test_painn_translation_and_wrap_invariance_full_size(painn_full):
    """A rigid shift by an arbitrary vector (atoms leave the cell: unwrapped positions) changes no distance."""
    shift = np.array([1.2345, -7.891, 0.377])
    moved = [dict(c, positions=c["positions"] + shift) for c in painn_full["chains"]]
    e, f, _, _ = _run(painn_full["eng"], moved)
    for k in range(C_FULL):
        assert abs(e[k] - painn_full["e"][k]) <= 2 * E_TOL_PER_ATOM * painn_full["n"][k], (k, e[k], painn_full["e"][k])
        assert np.abs(f[k] - painn_full["f"][k]).max() <= 2 * _ftol(painn_full["fmax"][k]), k


def test_painn_permutation_equivariance_full_size(painn_full):
    """Relabelling the atoms permutes the forces and leaves the energy alone (rows are receiver-sorted, so every
    per-receiver sum runs in a different order: an fp32 re-association, inside the tolerance band)."""
    rng = np.random.default_rng(5)
    perms = [rng.permutation(n) for n in painn_full["n"]]
    shuffled = [dict(c, positions=c["positions"][p], numbers=c["numbers"][p]) for c, p in zip(painn_full["chains"], perms)]
    e, f, _, _ = _run(painn_full["eng"], shuffled)
    for k in range(C_FULL):
        assert abs(e[k] - painn_full["e"][k]) <= 2 * E_TOL_PER_ATOM * painn_full["n"][k], k
        assert np.abs(f[k] - painn_full["f"][k][perms[k]]).max() <= 2 * _ftol(painn_full["fmax"][k]), k


def test_painn_rotation_equivariance_full_size(painn_full):
    """Rotating positions AND cell by a proper rotation (a generic axis: the cell is no longer axis-aligned) rotates the
    forces and keeps the energy: exercises the vector channel (v, U v, V v) and its hand-written backward."""
    ax = np.array([0.3, -0.5, 0.81])
    ax /= np.linalg.norm(ax)
    th = 0.7
    K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
    R = np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K
    rot = [dict(c, positions=c["positions"] @ R.T, cell=np.asarray(c["cell"]) @ R.T) for c in painn_full["chains"]]
    e, f, _, _ = _run(painn_full["eng"], rot)
    for k in range(C_FULL):
        assert abs(e[k] - painn_full["e"][k]) <= 2 * E_TOL_PER_ATOM * painn_full["n"][k], (k, e[k], painn_full["e"][k])
        assert np.abs(f[k] - painn_full["f"][k] @ R.T).max() <= 2 * _ftol(painn_full["fmax"][k]), k


def test_painn_memo_constrained_engine_equals_plain_full_size(painn_full, structures):
    """The engine as bench.py configures it (frozen-pair filter memo, group kernels, constrained gradients) against the
    plain one on the full batch: energies in the band, forces of the free atoms in the band, frozen atoms report 0."""
    from surface_sampling_b200 import engine
    s = structures["SrTiO3_001_2x2"]
    fixed0 = orelax.fixed_mask_from_surface_depth(s["positions"], s["cell"], 1)
    # the memo is keyed on the framework's coordinates: every chain carries the SAME (unperturbed) framework rows
    chains = [dict(c, positions=np.vstack([s["positions"], c["positions"][60:]])) for c in painn_full["chains"]]
    fix = [np.concatenate([fixed0, np.zeros(n - 60, bool)]) for n in painn_full["n"]]
    e0, f0, _, _ = _run(painn_full["eng"], chains, fix)
    memo = engine.PainnEngine(painn_full["states"], None)
    memo.set_framework(s["positions"], s["cell"], PBC3, fixed0, constrained_forces=True)
    b1 = _batch(chains, fix)
    r1 = memo.energy_forces(b1, constrained_forces=True)
    e1, f1 = r1["energy"].cpu().numpy().astype(np.float64), b1.split_host(r1["forces"].cpu().numpy().astype(np.float64))
    # raw-force mode of the same memoised engine: every atom's force, frozen ones included
    e2, f2, _, _ = _run(memo, chains, fix)
    for k in range(C_FULL):
        fm = np.abs(f0[k]).max()
        assert abs(e1[k] - e0[k]) <= 2 * E_TOL_PER_ATOM * painn_full["n"][k], (k, e1[k], e0[k])
        free = ~fix[k]
        assert np.abs(f1[k][free] - f0[k][free]).max() <= 2 * _ftol(fm), k
        assert not f1[k][~free].any()
        assert abs(e2[k] - e0[k]) <= 2 * E_TOL_PER_ATOM * painn_full["n"][k] and np.abs(f2[k] - f0[k]).max() <= 2 * _ftol(fm), k


def test_painn_duplicate_chains_bitwise_full_size(painn_full):
    """The same structure at different places of the full batch gives the same BITS (row scheduling, sender windows and
    group membership depend on a structure's own content only)."""
    chains = list(painn_full["chains"])
    chains[17], chains[64], chains[127] = chains[3], chains[3], chains[3]
    _, _, r, b = _run(painn_full["eng"], chains)
    ptr = b.atom_ptr_host
    ref_f = r["forces"][ptr[3]:ptr[4]]
    for k in (17, 64, 127):
        assert torch.equal(r["forces"][ptr[k]:ptr[k + 1]], ref_f) and r["energy"][k] == r["energy"][3], k
        assert torch.equal(r["forces_std"][ptr[k]:ptr[k + 1]], r["forces_std"][ptr[3]:ptr[4]])
    assert np.array_equal(r["energy"][:3].cpu().numpy(), painn_full["e"][:3])     # ... and of the rest of the batch


def test_nbr_list_full_batch_symmetric_sorted_complete(painn_full):
    """Full batch (11.8 k atoms, ~0.8 M edges at 6 A): every edge has its reverse with the opposite image shift, rows are
    sorted by (neighbour, shift), no pair is missed or invented on sampled rows (brute force over 27 images)."""
    from surface_sampling_b200 import engine
    chains = painn_full["chains"]
    b = engine.Batch.from_arrays([c["positions"] for c in chains], [c["numbers"] for c in chains],
                                 [c["cell"] for c in chains], [PBC3 for _ in chains])
    rowptr, col, shift = engine.neighbor_list(b, 6.0)
    torch.cuda.synchronize()
    rp, cj, sh = rowptr.cpu().numpy().astype(np.int64), col.cpu().numpy().astype(np.int64), shift.cpu().numpy()[:, :3].astype(np.int64)
    A = len(rp) - 1
    ci = np.repeat(np.arange(A), np.diff(rp))
    assert len(cj) == rp[-1] > 700_000
    key = lambda i, j, s: ((i * A + j) * 27 + (s[:, 0] + 1) * 9 + (s[:, 1] + 1) * 3 + (s[:, 2] + 1))
    assert np.abs(sh).max() <= 1
    fwd, rev = key(ci, cj, sh), key(cj, ci, -sh)
    assert len(np.unique(fwd)) == len(fwd)                      # no duplicates
    assert np.array_equal(np.sort(fwd), np.sort(rev))           # symmetric
    assert (np.diff(fwd)[np.diff(ci) == 0] > 0).all()           # receiver-sorted rows, ascending (j, shift) inside a row
    ptr = b.atom_ptr_host
    rng = np.random.default_rng(0)
    for k in rng.choice(C_FULL, 6, replace=False):
        pos, cell = chains[k]["positions"], np.asarray(chains[k]["cell"])
        imgs = np.array([[a, b_, c] for a in (-1, 0, 1) for b_ in (-1, 0, 1) for c in (-1, 0, 1)])
        for il in rng.choice(len(pos), 5, replace=False):
            d = pos[None, :, :] + (imgs @ cell)[:, None, :] - pos[il]
            dist = np.linalg.norm(d, axis=2)
            want = {(int(j), tuple(imgs[m])) for m, j in zip(*np.nonzero(dist < 6.0)) if not (j == il and not imgs[m].any())}
            # stay clear of the cutoff sphere's surface: there the fp64 comparison here and the kernel's own may differ
            shell = {(int(j), tuple(imgs[m])) for m, j in zip(*np.nonzero(np.abs(dist - 6.0) < 1e-9))}
            i = ptr[k] + il
            got = {(int(cj[q] - ptr[k]), tuple(sh[q])) for q in range(rp[i], rp[i + 1])}
            assert got - shell == want - shell, (k, il)


# ------------------------------------------------------------------ classical potentials ------------------------------

def _types(numbers, table):
    return np.array([table[int(z)] for z in numbers], dtype=np.int32)


@pytest.mark.parametrize("kind,n_chains", [("tersoff", 256), ("sw", 1024)])
def test_classical_properties_full_size(structures, potentials, kind, n_chains):
    """Config 2 / 3 sizes (256 GaN chains; 1024 Si chains = 128 per GPU x 8): fp64 kernels, so the properties hold to
    fp64 round-off -- net force, rigid translation, and bitwise equality of duplicated chains anywhere in the batch."""
    from surface_sampling_b200 import engine
    rng = np.random.default_rng(11)
    if kind == "tersoff":
        eng = engine.ClassicalEngine(engine.POT_TERSOFF, engine.tersoff_param_table(potentials["GaN.tersoff"], ["Ga", "N"]), 2,
                                     n_max=64, max_nbr=32)
        base, tab = structures["GaN_0001_3x3"], {31: 0, 7: 1}
    else:
        eng = engine.ClassicalEngine(engine.POT_SW, engine.sw_param_table(), 1, n_max=160, max_nbr=40)
        base, tab = structures["Si_111_5x5"], {14: 0}
    uniq = [perturbed(base, rng, 0.06) for _ in range(32)]
    structs = [uniq[k % 32] for k in range(n_chains)]
    types = [_types(s["numbers"], tab) for s in structs]

    def run(ss):
        b = engine.Batch.from_arrays([s["positions"] for s in ss], types, [s["cell"] for s in ss], [s["pbc"] for s in ss])
        r = eng.energy_forces(b)
        torch.cuda.synchronize()
        return r["energy"].cpu().numpy(), r["forces"].cpu().numpy().reshape(n_chains, -1, 3)

    e, f = run(structs)
    assert np.isfinite(e).all() and np.abs(f.sum(axis=1)).max() < 1e-9
    for k in range(32, n_chains):                      # duplicates of chain k % 32: same bits
        assert e[k] == e[k % 32]
    assert np.array_equal(f[32:64], f[:32]) and np.array_equal(f[-32:], f[:32])
    e2, f2 = run([dict(s, positions=s["positions"] + np.array([3.21, -1.07, 0.0])) for s in structs])
    assert np.abs(e2 - e).max() < 1e-9 * np.abs(e).max() and np.abs(f2 - f).max() < 1e-8
