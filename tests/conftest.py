"""Shared fixtures.  `-m gpu` tests need a B200 and call the CUDA path through the C ABI;
everything else runs on CPU (oracle vs golden vectors, host logic, symbol checks)."""
import json
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def structures():
    z = np.load(GOLDEN / "structures.npz")
    names = sorted({k.split("/")[0] for k in z.files})
    return {n: {f: z[f"{n}/{f}"] for f in ("numbers", "positions", "cell", "pbc", "fixed")} for n in names}


@pytest.fixture(scope="session")
def potentials():
    return json.loads((GOLDEN / "potentials.json").read_text())


@pytest.fixture(scope="session")
def golden_values():
    return json.loads((GOLDEN / "golden_values.json").read_text())


@pytest.fixture(scope="session")
def sto_weights():
    from oracle.painn import load_golden_weights
    return load_golden_weights(GOLDEN / "painn_sto_weights.npz")


def perturbed(s, rng, sigma=0.05):
    out = dict(s)
    out["positions"] = s["positions"] + rng.normal(0, sigma, s["positions"].shape)
    return out


def with_adsorbates(s, rng, n_ads, species, height=1.5):
    """Append n_ads atoms above random top-layer atoms (virtual-site style insertions)."""
    pos, num = s["positions"], s["numbers"]
    top = np.argsort(pos[:, 2])[-8:]
    new_p, new_z = [], []
    for _ in range(n_ads):
        a = rng.choice(top)
        new_p.append(pos[a] + np.array([rng.uniform(-1, 1), rng.uniform(-1, 1), height]))
        new_z.append(rng.choice(species))
    out = dict(s)
    out["positions"] = np.concatenate([pos, np.array(new_p, dtype=float).reshape(-1, 3)])
    out["numbers"] = np.concatenate([num, np.array(new_z, dtype=num.dtype)])
    return out
