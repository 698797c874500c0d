"""Model / data loaders of the drop-in: what ``scripts/sample_surface.py:154-175`` does before it builds the
calculator, without NFF / ASE / CatKit installed.

  load_model(path, model_type="PaiNN")   ~ nff.train.builders.model.load_model  (scripts/sample_surface.py:166):
        reads a ``best_model`` torch pickle (or the folder that holds it) and returns the PaiNN state dict
        (checkpoint keys) as fp32 numpy arrays — what ``EnsembleNFF`` packs for the GPU.
  as_state_dict(model)                   any of: state dict, ``torch.nn.Module`` (real NFF ``Painn``), path.
  init_random_weights(seed)              random-init PaiNN of the checkpoint's shapes (BASELINE.json: "random-init
        PaiNN weights (Zenodo checkpoints are unavailable offline)").
  load_slab_pickle(path)                 the pristine-slab pickles (``pickle.load`` at scripts/sample_surface.py:155-157)
        -> shim ``Atoms`` with its FixAtoms.
  load_offset_data(path)                 offset_data.json (scripts/sample_surface.py:131-143).

The ``best_model`` pickles reference 14 NFF classes by module path; when NFF is not importable they are served by
empty ``nn.Module`` subclasses registered under those paths for the duration of the load (the state dict is all the
engine needs: architecture constants are fixed by csrc/painn_layout.h and checked against the tensor shapes).
"""
from __future__ import annotations

import json
import math
import pickle
import sys
import types
from pathlib import Path

import numpy as np

from .engine import F, F3, FH, NCONV, NEMB, NRBF

_NFF_CLASSES = {
    "nff.nn.activations": ["Swish"],
    "nff.nn.layers": ["CosineEnvelope", "Dense", "PainnRadialBasis"],
    "nff.nn.models.painn": ["Painn"],
    "nff.nn.modules.painn": ["DistanceEmbed", "EmbeddingBlock", "InvariantDense", "InvariantMessage", "MessageBlock",
                             "ReadoutBlock", "UpdateBlock"],
    "nff.nn.modules.schnet": ["ScaleShift", "SumPool"],
}


def expected_shapes() -> dict:
    """State-dict keys and shapes of the PaiNN the engine implements."""
    sh = {"embed_block.atom_embed.weight": (NEMB, F)}
    for l in range(NCONV):
        p, u = f"message_blocks.{l}.inv_message.", f"update_blocks.{l}."
        sh.update({p + "inv_dense.layers.0.weight": (F, F), p + "inv_dense.layers.0.bias": (F,),
                   p + "inv_dense.layers.1.weight": (F3, F), p + "inv_dense.layers.1.bias": (F3,),
                   p + "dist_embed.block.1.weight": (F3, NRBF), p + "dist_embed.block.1.bias": (F3,),
                   u + "u_mat.weight": (F, F), u + "v_mat.weight": (F, F),
                   u + "s_dense.0.weight": (F, 2 * F), u + "s_dense.0.bias": (F,),
                   u + "s_dense.1.weight": (F3, F), u + "s_dense.1.bias": (F3,)})
    r = "readout_blocks.0.readoutdict.energy."
    sh.update({r + "0.weight": (FH, F), r + "0.bias": (FH,), r + "1.weight": (1, FH), r + "1.bias": (1,)})
    return sh


def check_state_dict(sd: dict) -> dict:
    """Validate keys/shapes against the architecture the kernels are compiled for; returns fp32 numpy arrays."""
    out = {}
    for k, shape in expected_shapes().items():
        if k not in sd:
            raise KeyError(f"PaiNN state dict lacks {k!r}: not the feat-128 / 3-conv / 20-rbf PaiNN the engine serves")
        a = sd[k]
        a = a.detach().cpu().numpy() if hasattr(a, "detach") else np.asarray(a)
        if tuple(a.shape) != shape:
            raise ValueError(f"{k}: shape {tuple(a.shape)} != {shape}")
        out[k] = np.ascontiguousarray(a, dtype=np.float32)
    return out


class _StubModules:
    """Temporarily registers empty stand-ins for the NFF classes a ``best_model`` pickle names."""

    def __enter__(self):
        import torch.nn as nn
        self.added = []
        try:
            import nff.nn.models.painn  # noqa: F401  (the real package wins when present)
            return self
        except Exception:
            pass
        for modname, classes in _NFF_CLASSES.items():
            parts = modname.split(".")
            for i in range(1, len(parts) + 1):
                name = ".".join(parts[:i])
                if name not in sys.modules:
                    sys.modules[name] = types.ModuleType(name)
                    self.added.append(name)
            mod = sys.modules[modname]
            for cls in classes:
                if not hasattr(mod, cls):
                    base = nn.Linear if cls == "Dense" else nn.Module
                    setattr(mod, cls, type(cls, (base,), {"__module__": modname}))
        return self

    def __exit__(self, *exc):
        for name in self.added:
            sys.modules.pop(name, None)
        return False


def load_model(path, model_type: str = "PaiNN", map_location="cpu", **kwargs) -> dict:
    """``load_model(model_path, model_type=model_type, map_location=device)`` of the reference scripts.
    `path` is the ``best_model`` file or the folder containing it.  Only PaiNN is served by the B200 engine
    (the reference's CHGNet / NffScaleMACE branches are other model families, outside the hot path)."""
    if model_type not in ("PaiNN", "Painn", "painn"):
        raise NotImplementedError(f"model_type {model_type!r}: the B200 engine implements PaiNN only")
    import torch
    p = Path(path)
    if p.is_dir():
        p = p / "best_model"
    if not p.exists():
        raise FileNotFoundError(p)
    with _StubModules():
        model = torch.load(p, map_location="cpu", weights_only=False)
    sd = model.state_dict() if hasattr(model, "state_dict") else model
    return check_state_dict(sd)


def as_state_dict(model) -> dict:
    """Accept what ``EnsembleNFF(models, ...)`` may be handed: a state dict, a torch module, or a checkpoint path."""
    if isinstance(model, (str, Path)):
        return load_model(model)
    if isinstance(model, dict):
        return check_state_dict(model)
    if hasattr(model, "state_dict"):
        return check_state_dict(model.state_dict())
    raise TypeError(f"cannot interpret {type(model).__name__} as a PaiNN model")


def init_random_weights(seed: int) -> dict:
    """Random-init PaiNN with the checkpoint's shapes: Xavier-uniform matrices, zero biases, N(0,1) embedding with
    padding row 0 (what ``nn.Embedding(padding_idx=0)`` / NFF ``Dense`` give a freshly built model)."""
    import torch
    g = torch.Generator().manual_seed(int(seed))
    sd = {}
    for k, shape in expected_shapes().items():
        if k == "embed_block.atom_embed.weight":
            a = torch.randn(*shape, generator=g).numpy().astype(np.float32)
            a[0] = 0
        elif len(shape) == 2:
            bound = math.sqrt(6.0 / (shape[0] + shape[1]))
            a = ((torch.rand(*shape, generator=g) * 2 - 1) * bound).numpy().astype(np.float32)
        else:
            a = np.zeros(shape, np.float32)
        sd[k] = a
    return sd


# ------------------------------------------------------------------------------------------------
class _Stub:
    def __init__(self, *a, **k):
        pass

    def __setstate__(self, state):
        if isinstance(state, dict):
            self.__dict__.update(state)
        else:
            self.__dict__["_state"] = state


class _SlabUnpickler(pickle.Unpickler):
    """ase.Atoms / catkit Gratoms pickles written under numpy 1.x, read without ASE."""

    def find_class(self, module, name):
        if module.startswith("numpy"):
            return super().find_class(module.replace("numpy.core", "numpy._core"), name)
        if module in ("builtins", "collections", "copyreg"):
            return super().find_class(module, name)
        return type(name, (_Stub,), {})


def load_slab_pickle(path):
    """Pristine-slab pickle -> shim Atoms (numbers, positions, cell, pbc, FixAtoms, tags / surface_atoms arrays)."""
    from .atoms import Atoms, FixAtoms
    try:
        import ase  # noqa: F401
        with open(path, "rb") as f:
            return pickle.load(f)          # real ASE present: hand the object over unchanged
    except ImportError:
        pass
    with open(path, "rb") as f:
        d = _SlabUnpickler(f).load().__dict__
    arrays = d["arrays"]
    a = Atoms(numbers=np.array(arrays["numbers"], dtype=int), positions=np.array(arrays["positions"], dtype=float),
              cell=np.array(d["_cellobj"].__dict__["array"], dtype=float), pbc=np.array(d["_pbc"], dtype=bool))
    cons = d.get("_constraints", [])
    if cons and cons[0].__dict__.get("index", None) is not None:
        a.set_constraint(FixAtoms(indices=np.array(cons[0].__dict__["index"], dtype=int)))
    for k in ("tags", "surface_atoms"):
        if k in arrays:
            a.set_array(k, np.array(arrays[k]))
    return a


def load_eam_funcfl(path) -> dict:
    """LAMMPS single-element `funcfl` EAM file (mcmc/potentials/{Cu,Au}_u3.eam):
    line 2 = Z mass a0 lattice, line 3 = Nrho drho Nr dr rc, then F(rho)[Nrho], Z(r)[Nr], rho(r)[Nr]."""
    lines = Path(path).read_text().splitlines()
    head, grid = lines[1].split(), lines[2].split()
    nrho, drho, nr, dr, rc = int(grid[0]), float(grid[1]), int(grid[2]), float(grid[3]), float(grid[4])
    vals = np.array(" ".join(lines[3:]).split(), dtype=np.float64)
    if vals.size != nrho + 2 * nr:
        raise ValueError(f"{path}: expected {nrho + 2 * nr} table values, found {vals.size}")
    return {"Z": int(head[0]), "mass": float(head[1]), "a0": float(head[2]), "nrho": nrho, "drho": drho, "nr": nr,
            "dr": dr, "rc": rc, "frho": vals[:nrho], "zr": vals[nrho:nrho + nr], "rhor": vals[nrho + nr:]}


def load_offset_data(path) -> dict:
    return json.loads(Path(path).read_text())
