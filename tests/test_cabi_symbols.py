"""The C-ABI library loads on a CPU-only box and exports every symbol include/vssr_b200.h declares
(no compute calls here)."""
import re
from pathlib import Path

from surface_sampling_b200 import _lib, engine

ROOT = Path(__file__).resolve().parent.parent


def header_symbols():
    text = (ROOT / "include" / "vssr_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vssr_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    syms = header_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in the header but not exported"
    assert sorted(_lib.SIGNATURES) == syms, "ctypes table and header disagree"


def test_host_only_queries():
    lib = _lib.load()
    assert lib.vssr_version() == 100
    assert int(lib.vssr_painn_weight_floats()) == engine.painn_weight_floats()
    assert lib.vssr_painn_workspace_bytes(3, 60, 4000) > 3 * 60 * 128 * 4
    assert lib.vssr_classical_smem_bytes(0, 64, 24) < 227 * 1024


def test_weight_packing_roundtrip():
    import numpy as np
    from oracle.painn import init_random_weights
    sd = init_random_weights(0)
    flat = engine.pack_painn_weights(sd)
    assert flat.dtype == np.float32 and flat.size == engine.painn_weight_floats()
    # embedding first, W1T of layer 0 right after
    assert np.array_equal(flat[:100 * 128].reshape(100, 128), sd["embed_block.atom_embed.weight"])
    w1 = sd["message_blocks.0.inv_message.inv_dense.layers.0.weight"]
    assert np.array_equal(flat[100 * 128:100 * 128 + 128 * 128].reshape(128, 128), w1.T)
