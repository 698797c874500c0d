"""Max |dE|/atom and |dF| of the CUDA path vs the fp64 oracle on sane structures (run on the GPU box).
usage: [VSSR_GEMM=fma] python profiles/error_probe.py"""
import json, sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle.painn import EnsembleOracle, load_golden_weights
from surface_sampling_b200 import engine
z = np.load(ROOT / "tests/golden/structures.npz")
pots = json.loads((ROOT / "tests/golden/potentials.json").read_text())
ws = load_golden_weights(ROOT / "tests/golden/painn_sto_weights.npz")
eng = engine.PainnEngine(ws, pots["offset_data"])
ens = EnsembleOracle(ws, pots["offset_data"], dtype=torch.float64)
rng = np.random.default_rng(0)
pbc = np.array([True] * 3)
structs = []
for n in ("SrTiO3_001_2x2", "O44Sr12Ti16", "O40Sr16Ti12"):
    p = z[f"{n}/positions"]
    for s in (0.0, 0.03, 0.08):
        structs.append((p + rng.normal(0, s, p.shape), z[f"{n}/numbers"], z[f"{n}/cell"]))
b = engine.Batch.from_arrays([s[0] for s in structs], [s[1] for s in structs], [s[2] for s in structs], [pbc] * len(structs))
r = eng.energy_forces(b)
e = r["energy"].cpu().numpy(); f = b.split_host(r["forces"].cpu().numpy())
de, df, fmax = 0, 0, 0
for k, s in enumerate(structs):
    o = ens.calculate(s[0], s[1], s[2], pbc)
    de = max(de, abs(e[k] - o["energy"][0]) / len(s[1])); df = max(df, np.abs(f[k] - o["forces"]).max()); fmax = max(fmax, np.abs(o["forces"]).max())
print("max |dE|/atom = %.3e eV (tol 1e-5), max |dF| = %.3e eV/A (tol 1e-4), max |F| = %.2f" % (de, df, fmax))
