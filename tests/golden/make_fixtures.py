"""Generate the committed golden fixtures from the reference DATA files.

Run HERE (the build container) only: ``python tests/golden/make_fixtures.py``.
``/root/reference`` does not exist on the GPU box, so everything the tests, ``smoke()``
and ``bench.py`` need is converted into plain ``.npz``/``.json`` files next to this script.

What is read (data, never source code):
  * slab pickles (``catkit.gratoms.Gratoms`` pickled under numpy 1.x)   SURVEY.md App. B.2
  * P1 CIFs written by ASE                                               SURVEY.md App. B.3
  * PaiNN ``best_model`` torch zip-pickles (state_dict only is kept)     SURVEY.md App. B.1
  * ``GaN.tersoff`` / ``*_u3.eam`` potential tables                      SURVEY.md App. B.4
  * ``offset_data.json``
No third-party package (ase, nff, catkit) is needed: stub classes absorb the pickled objects.
"""
from __future__ import annotations

import io
import json
import pickle
import sys
import types
from pathlib import Path

import numpy as np

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent


# --------------------------------------------------------------------------------------
# slab pickles
# --------------------------------------------------------------------------------------
class _Stub:
    """Permissive stand-in for any non-numpy class found in a pickle."""

    def __init__(self, *a, **k):
        pass

    def __setstate__(self, state):
        if isinstance(state, dict):
            self.__dict__.update(state)
        else:
            self.__dict__["_state"] = state


class _SlabUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module.startswith("numpy"):
            module = module.replace("numpy.core", "numpy._core")
            return super().find_class(module, name)
        if module in ("builtins", "collections", "copyreg"):
            return super().find_class(module, name)
        return type(name, (_Stub,), {})


def load_slab_pickle(path: Path) -> dict:
    with open(path, "rb") as f:
        obj = _SlabUnpickler(f).load()
    d = obj.__dict__
    arrays = d["arrays"]
    cellobj = d["_cellobj"]
    cell = np.array(cellobj.__dict__["array"], dtype=np.float64)
    pbc = np.array(d["_pbc"], dtype=bool)
    fixed = np.zeros(0, dtype=np.int64)
    cons = d.get("_constraints", [])
    if cons:
        idx = cons[0].__dict__.get("index", None)
        if idx is not None:
            fixed = np.array(idx, dtype=np.int64)
    return {
        "numbers": np.array(arrays["numbers"], dtype=np.int64),
        "positions": np.array(arrays["positions"], dtype=np.float64),
        "cell": cell,
        "pbc": pbc,
        "fixed": fixed,
    }


# --------------------------------------------------------------------------------------
# CIF (ASE-written P1)
# --------------------------------------------------------------------------------------
_ATOM_COLS = ["_atom_site_type_symbol", "_atom_site_label", "_atom_site_symmetry_multiplicity",
              "_atom_site_fract_x", "_atom_site_fract_y", "_atom_site_fract_z",
              "_atom_site_occupancy"]
_Z = {"H": 1, "N": 7, "O": 8, "Si": 14, "Ti": 22, "Cu": 29, "Ga": 31, "Sr": 38, "Au": 79}


def load_cif(path: Path) -> dict:
    a = b = c = al = be = ga = None
    rows = []
    cols: list[str] = []
    in_loop = False
    for line in path.read_text().splitlines():
        t = line.split()
        if not t:
            continue
        if t[0] == "_cell_length_a":
            a = float(t[1])
        elif t[0] == "_cell_length_b":
            b = float(t[1])
        elif t[0] == "_cell_length_c":
            c = float(t[1])
        elif t[0] == "_cell_angle_alpha":
            al = float(t[1])
        elif t[0] == "_cell_angle_beta":
            be = float(t[1])
        elif t[0] == "_cell_angle_gamma":
            ga = float(t[1])
        elif t[0] == "loop_":
            in_loop = True
            cols = []
        elif in_loop and t[0].startswith("_"):
            cols.append(t[0])
        elif in_loop and "_atom_site_fract_x" in cols and len(t) == len(cols):
            rows.append(t)
    # ASE cell convention: a along x, b in xy plane
    al_r, be_r, ga_r = np.deg2rad([al, be, ga])
    va = np.array([a, 0.0, 0.0])
    vb = np.array([b * np.cos(ga_r), b * np.sin(ga_r), 0.0])
    cx = c * np.cos(be_r)
    cy = c * (np.cos(al_r) - np.cos(be_r) * np.cos(ga_r)) / np.sin(ga_r)
    cz = np.sqrt(max(c * c - cx * cx - cy * cy, 0.0))
    cell = np.array([va, vb, [cx, cy, cz]])
    cell[np.abs(cell) < 1e-12] = 0.0
    numbers, frac = [], []
    ci = {name: k for k, name in enumerate(_ATOM_COLS)}
    for t in rows:
        numbers.append(_Z[t[ci["_atom_site_type_symbol"]]])
        frac.append([float(t[ci["_atom_site_fract_x"]]), float(t[ci["_atom_site_fract_y"]]),
                     float(t[ci["_atom_site_fract_z"]])])
    frac = np.array(frac)
    return {
        "numbers": np.array(numbers, dtype=np.int64),
        "positions": frac @ cell,
        "cell": cell,
        "pbc": np.array([True, True, True]),
        "fixed": np.zeros(0, dtype=np.int64),
    }


def _isfloat(x):
    try:
        float(x)
        return True
    except ValueError:
        return False


# --------------------------------------------------------------------------------------
# PaiNN checkpoints -> state dict
# --------------------------------------------------------------------------------------
_NFF_CLASSES = {
    "nff.nn.activations": ["Swish"],
    "nff.nn.layers": ["CosineEnvelope", "Dense", "PainnRadialBasis"],
    "nff.nn.models.painn": ["Painn"],
    "nff.nn.modules.painn": [
        "DistanceEmbed", "EmbeddingBlock", "InvariantDense", "InvariantMessage",
        "MessageBlock", "ReadoutBlock", "UpdateBlock",
    ],
    "nff.nn.modules.schnet": ["ScaleShift", "SumPool"],
}


def _install_nff_stubs():
    import torch.nn as nn

    for modname, classes in _NFF_CLASSES.items():
        parts = modname.split(".")
        for i in range(1, len(parts) + 1):
            name = ".".join(parts[:i])
            if name not in sys.modules:
                sys.modules[name] = types.ModuleType(name)
        mod = sys.modules[modname]
        for cls in classes:
            base = nn.Linear if cls == "Dense" else nn.Module
            setattr(mod, cls, type(cls, (base,), {"__module__": modname}))


def load_painn_state(path: Path) -> dict:
    import torch

    _install_nff_stubs()
    model = torch.load(path, map_location="cpu", weights_only=False)
    sd = model.state_dict()
    return {k: v.detach().cpu().numpy().astype(np.float32) for k, v in sd.items()}


# --------------------------------------------------------------------------------------
# potentials
# --------------------------------------------------------------------------------------
def load_tersoff(path: Path) -> dict:
    toks = []
    for line in path.read_text().splitlines():
        line = line.split("#")[0]
        toks += line.split()
    entries = []
    for i in range(0, len(toks), 17):
        e = toks[i:i + 17]
        entries.append({"elements": e[:3], "params": [float(x) for x in e[3:]]})
    return {
        "order": ["m", "gamma", "lambda3", "c", "d", "costheta0", "n", "beta", "lambda2",
                  "B", "R", "D", "lambda1", "A"],
        "entries": entries,
    }


def load_funcfl(path: Path) -> dict:
    lines = path.read_text().splitlines()
    t2 = lines[1].split()
    t3 = lines[2].split()
    nrho, drho, nr, dr, rc = int(t3[0]), float(t3[1]), int(t3[2]), float(t3[3]), float(t3[4])
    vals = np.array(" ".join(lines[3:]).split(), dtype=np.float64)
    assert vals.size == nrho + 2 * nr, (vals.size, nrho, nr)
    return {
        "Z": int(t2[0]), "mass": float(t2[1]), "a0": float(t2[2]),
        "nrho": nrho, "drho": drho, "nr": nr, "dr": dr, "rc": rc,
        "frho": vals[:nrho], "zr": vals[nrho:nrho + nr], "rhor": vals[nrho + nr:],
    }


def main():
    slabs = {
        "GaN_0001_3x3": REF / "tutorials/data/GaN_0001/GaN_0001_3x3_pristine_slab.pkl",
        "Si_111_5x5": REF / "tutorials/data/Si_111_5x5/Si_111_5x5_pristine_slab.pkl",
        "SrTiO3_001_2x2": REF / "tutorials/data/SrTiO3_001/SrTiO3_001_2x2_pristine_slab.pkl",
        "SrTiO3_001_2x2x4": REF / "tutorials/data/SrTiO3_001/SrTiO3_001_2x2x4_pristine_slab.pkl",
        "Au_110_2x2": REF / "tests/data/Au_110/Au_110_2x2_pristine_slab.pkl",
        "SrTiO3_unit_cell": REF / "tests/data/SrTiO3_001/SrTiO3_unit_cell.pkl",
    }
    cifs = {
        "O44Sr12Ti16": REF / "tests/data/SrTiO3_001/O44Sr12Ti16.cif",
        "O36Sr12Ti12": REF / "tests/data/SrTiO3_001/O36Sr12Ti12.cif",
        "O40Sr16Ti12": REF / "tests/data/SrTiO3_001/O40Sr16Ti12.cif",
        "SrTiO3_001_distance_failed": REF / "tests/data/SrTiO3_001/SrTiO3_001_distance_failed.cif",
        "Au_110_2x2_proper_adsorbed": REF / "tests/data/Au_110/Au_110_2x2_proper_adsorbed_slab.cif",
    }
    out = {}
    for name, p in slabs.items():
        s = load_slab_pickle(p)
        for k, v in s.items():
            out[f"{name}/{k}"] = v
        print(name, len(s["numbers"]), "atoms, fixed", len(s["fixed"]), "pbc", s["pbc"])
    for name, p in cifs.items():
        s = load_cif(p)
        for k, v in s.items():
            out[f"{name}/{k}"] = v
        print(name, len(s["numbers"]), "atoms (cif)")
    np.savez_compressed(OUT / "structures.npz", **out)

    w = {}
    for m in ("model01", "model02", "model03"):
        sd = load_painn_state(REF / f"tutorials/data/SrTiO3_001/nff/{m}/best_model")
        n = sum(v.size for v in sd.values())
        print(m, len(sd), "tensors", n, "params")
        for k, v in sd.items():
            w[f"{m}/{k}"] = v
    np.savez_compressed(OUT / "painn_sto_weights.npz", **w)

    pots = {
        "GaN.tersoff": load_tersoff(REF / "mcmc/potentials/GaN.tersoff"),
        "offset_data": json.loads((REF / "tutorials/data/SrTiO3_001/nff/offset_data.json").read_text()),
    }
    (OUT / "potentials.json").write_text(json.dumps(pots, indent=1))
    eam = {}
    for el in ("Cu", "Au"):
        f = load_funcfl(REF / f"mcmc/potentials/{el}_u3.eam")
        for k, v in f.items():
            eam[f"{el}/{k}"] = np.asarray(v)
    np.savez_compressed(OUT / "eam_funcfl.npz", **eam)
    print("written to", OUT)


if __name__ == "__main__":
    main()
