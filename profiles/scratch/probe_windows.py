"""Disambiguate the `huge` force mismatch: z-image interactions vs 3 sender windows."""
import sys, json, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from oracle.painn import EnsembleOracle, init_random_weights
from surface_sampling_b200 import engine
from test_gpu_painn import _stacked, _batch, PBC3
z = np.load('tests/golden/structures.npz')
S = lambda n: {f: z[f"{n}/{f}"] for f in ("numbers", "positions", "cell", "pbc", "fixed")}
states = [init_random_weights(s) for s in (0, 1, 2)]
eng = engine.PainnEngine(states, None)
ens = EnsembleOracle(states, None, dtype=torch.float64)
base, thick = S("SrTiO3_001_2x2"), S("SrTiO3_001_2x2x4")
def tall(s, c):
    s = dict(s); cell = s["cell"].copy(); cell[2, 2] = c; s["cell"] = cell; return s
cases = {
  "thin60_zimage": tall(base, 12.0),                       # 60 atoms, top/bottom 2.2 A apart through z
  "huge192_zimage": _stacked(thick, 112),                  # the failing case
  "huge192_noz": _stacked(tall(thick, 45.0), 112),         # same atoms, no z interaction
  "mid140_noz": _stacked(tall(thick, 45.0), 60),           # 140 atoms: 2 windows
  "big176_noz": _stacked(tall(thick, 45.0), 96),           # 176 atoms: 3 windows (bwd), 2 fwd
}
for name, s in cases.items():
    b = _batch([s]); r = eng.energy_forces(b)
    o = ens.calculate(s["positions"], s["numbers"], s["cell"], PBC3)
    f = r["forces"].cpu().numpy(); df = np.abs(f - o["forces"])
    bad = np.argsort(-df.max(1))[:5]
    print(name, len(s["numbers"]), "dE/atom %.2e" % (abs(r["energy"][0].item() - o["energy"][0]) / len(s["numbers"])),
          "max dF %.2e" % df.max(), "fscale %.2f" % np.abs(o["grads_per_model"]).max(), "worst atoms", bad.tolist(),
          [float("%.1e" % df[a].max()) for a in bad])
