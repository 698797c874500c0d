"""Timing of one ensemble evaluation with / without the frozen-pair filter memo (run on the GPU box)."""
import json, sys, time
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle.painn import init_random_weights
from surface_sampling_b200 import engine
z = np.load(ROOT / "tests/golden/structures.npz")
n = "SrTiO3_001_2x2"
pos, num, cell = z[f"{n}/positions"], z[f"{n}/numbers"], z[f"{n}/cell"]
pbc = np.array([True] * 3)
fixed = np.ones(60, bool); fixed[[7, 8, 22, 23, 37, 38, 52, 53]] = False
ws = [init_random_weights(s) for s in (0, 1, 2)]
C = 128
b = engine.Batch.from_arrays([pos] * C, [num] * C, [cell] * C, [pbc] * C, [fixed] * C)
for memo in (False, True):
    eng = engine.PainnEngine(ws, None)
    if memo:
        print("nslots", eng.set_framework(pos, cell, pbc, fixed))
    nb = engine.neighbor_list(b, 6.0)
    for _ in range(3):
        r = eng.energy_forces(b, z_host=np.concatenate([num] * C), nbrs=nb)
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(10):
        r = eng.energy_forces(b, z_host=np.concatenate([num] * C), nbrs=nb)
    torch.cuda.synchronize()
    print("memo" if memo else "plain", "%.3f ms/eval" % ((time.perf_counter() - t) * 100), float(r["energy"][0]))
