"""One ensemble evaluation of a 92-atom structure (two sender windows in the backward) + a 3-step relaxation with the
framework memo: the smallest run that touches every PaiNN kernel; meant for `compute-sanitizer --tool racecheck`."""
import sys
from pathlib import Path
import numpy as np
import torch
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from surface_sampling_b200 import engine, loaders
z = np.load(ROOT / "tests/golden/structures.npz")
base = {f: z[f"SrTiO3_001_2x2/{f}"] for f in ("numbers", "positions", "cell", "pbc", "fixed")}
ztop = base["positions"][:, 2].max()
grid = np.array([[0.5 + 1.9 * (a % 4) + 0.9 * ((a // 16) % 2), 0.5 + 1.9 * ((a // 4) % 4) + 0.9 * ((a // 16) % 2),
                  ztop + 1.5 + 1.9 * (a // 16)] for a in range(32)])
pos = np.vstack([base["positions"], grid])
num = np.concatenate([base["numbers"], np.array(([8, 38, 22, 8] * 8)[:32])])
PBC3 = np.array([True, True, True])
states = [loaders.init_random_weights(s) for s in (0, 1, 2)]
eng = engine.PainnEngine(states, None)
fixed0 = base["positions"][:, 2] < np.sort(base["positions"][:, 2])[-9]
fix = np.concatenate([fixed0, np.zeros(32, bool)])
b = engine.Batch.from_arrays([pos, base["positions"]], [num, base["numbers"]], [base["cell"]] * 2, [PBC3] * 2, [fix, fixed0])
r = eng.energy_forces(b)
torch.cuda.synchronize()
print("plain", r["energy"].cpu().numpy())
eng.set_framework(base["positions"], base["cell"], PBC3, fixed0, constrained_forces=True)
b = engine.Batch.from_arrays([pos, base["positions"]], [num, base["numbers"]], [base["cell"]] * 2, [PBC3] * 2, [fix, fixed0])
out = eng.relax(b, relax_steps=3, check=True)["out"].cpu().numpy()
print("relax", out[:, 0])
