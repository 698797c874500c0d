"""One batched FIRE relaxation (128 chains, SrTiO3(001) 2x2 + 0..5 adsorbates, 3-model PaiNN ensemble,
frozen-pair memo, constrained gradients) = the bench's device-resident step, for ncu.
usage: python profiles/relax_probe.py [n_relax]"""
import sys, time
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle.painn import init_random_weights          # weight init only
from surface_sampling_b200 import engine
z = np.load(ROOT / "tests/golden/structures.npz")
n = "SrTiO3_001_2x2"
pos, num, cell = z[f"{n}/positions"], z[f"{n}/numbers"], z[f"{n}/cell"]
pbc = np.array([True] * 3)
fixed = np.ones(60, bool); fixed[[7, 8, 22, 23, 37, 38, 52, 53]] = False
rng = np.random.default_rng(0)
C = 128
pl, nl, fl = [], [], []
for c in range(C):
    k = c % 6
    ads = np.column_stack([rng.uniform(0, 7.8, k), rng.uniform(0, 7.8, k), np.full(k, pos[:, 2].max() + 1.5)])
    pl.append(np.vstack([pos, ads])); nl.append(np.concatenate([num, rng.choice([8, 22, 38], k)]).astype(np.int64))
    fl.append(np.concatenate([fixed, np.zeros(k, bool)]))
eng = engine.PainnEngine([init_random_weights(s) for s in (0, 1, 2)], None)
eng.set_framework(pos, cell, pbc, fixed, constrained_forces=True)
for r in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    b = engine.Batch.from_arrays(pl, nl, [cell] * C, [pbc] * C, fl)
    torch.cuda.synchronize(); t = time.perf_counter()
    out = eng.relax(b, relax_steps=20, fmax=0.01, z_host=np.concatenate(nl), want_std=False, e_cap=C * 72 * 96)["out"].cpu()
    print("relax %d: %.2f ms  E0=%.6f" % (r, (time.perf_counter() - t) * 1e3, float(out[0, 0])))
