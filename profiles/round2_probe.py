"""One batched FIRE relaxation of the BURNT-IN bench regime for ncu: 128 chains of SrTiO3(001) 2x2 carrying 28..36
adsorbates on the bench's 64-site grid (the coverage bench.py reaches after its 200-step burn-in: mean 32.5), 3-model
PaiNN ensemble (random init), frozen-pair filter memo, constrained gradients.
usage: python profiles/round2_probe.py [n_relax] [relax_steps]"""
import sys, time
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from surface_sampling_b200 import engine, loaders, mc
z = np.load(ROOT / "tests/golden/structures.npz")
n = "SrTiO3_001_2x2"
pos, num, cell = z[f"{n}/positions"], z[f"{n}/numbers"], z[f"{n}/cell"]
pbc = np.array([True] * 3)
fixed = np.ones(60, bool); fixed[[7, 8, 22, 23, 37, 38, 52, 53]] = False
sites = mc.make_site_grid(pos, cell, 64, 1.5)
C = 128
pl, nl, fl = [], [], []
for c in range(C):
    rng = np.random.RandomState(c)
    k = 28 + c % 9
    pick = rng.choice(64, size=k, replace=False)
    pl.append(np.vstack([pos, sites[pick]])); nl.append(np.concatenate([num, rng.choice([8, 22, 38], k)]).astype(np.int64))
    fl.append(np.concatenate([fixed, np.zeros(k, bool)]))
eng = engine.PainnEngine([loaders.init_random_weights(s) for s in (0, 1, 2)], None, edges_per_atom=192)
eng.set_framework(pos, cell, pbc, fixed, constrained_forces=True)
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
for r in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    b = engine.Batch.from_arrays(pl, nl, [cell] * C, [pbc] * C, fl)
    torch.cuda.synchronize(); t = time.perf_counter()
    res = eng.relax(b, relax_steps=steps, fmax=0.01, z_host=np.concatenate(nl), want_std=False)
    out = res["out"].cpu()
    print("relax %d: %.2f ms  E0=%.6f status=%d atoms=%d" % (r, (time.perf_counter() - t) * 1e3, float(out[0, 0]), int(res["status"].item()), b.n_atoms))
print(eng.last_relax_edge_stats(b))
