"""ctypes binding of libvssr_b200.so (C ABI declared in include/vssr_b200.h).

The product path has NO CPU fallback: if the shared library is missing or a CUDA device is
required and absent, the calls raise.  The library is built in-tree by ``__graft_entry__.build()``
(``make -C surface_sampling_b200/csrc``).
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libvssr_b200.so"

_lib = None

c_void_p, c_int, c_i64, c_size_t, c_float, c_double = (C.c_void_p, C.c_int32, C.c_int64, C.c_size_t,
                                                       C.c_float, C.c_double)

# name -> (restype, argtypes); every symbol the header declares
SIGNATURES = {
    "vssr_version": (c_int, []),
    "vssr_device_cc": (c_int, []),
    "vssr_launch_count": (c_i64, []),
    "vssr_graph_launch_count": (c_i64, []),
    "vssr_kernel_class_count": (c_int, []),
    "vssr_profile_enable": (c_int, [c_int]),
    "vssr_profile_collect": (c_int, [c_void_p, c_void_p, c_int]),
    "vssr_nbr_build": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_void_p,
                               c_void_p, c_void_p, c_void_p, c_i64, c_void_p, c_void_p]),
    "vssr_painn_weight_floats": (c_i64, []),
    "vssr_painn_workspace_bytes": (c_size_t, [c_int, c_int, c_i64]),
    "vssr_painn_filter_cache_bytes": (c_size_t, [c_int, c_int, c_i64]),
    "vssr_painn_filter_cache_workspace_bytes": (c_size_t, [c_int, c_i64]),
    "vssr_painn_filter_cache_build": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_float,
                                              c_float, c_i64, c_void_p, c_size_t, c_void_p, c_size_t, c_void_p, c_void_p]),
    "vssr_painn_energy_grad": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                       c_void_p, c_void_p, c_void_p, c_i64, c_float, c_void_p, c_int, c_i64, c_int, c_void_p, c_size_t,
                                       c_void_p, c_void_p, c_void_p, c_void_p]),
    "vssr_ensemble_stats": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p,
                                    c_void_p, c_void_p, c_void_p, c_void_p]),
    "vssr_system_reduce": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "vssr_atom_norm": (c_int, [c_void_p, c_int, c_void_p, c_void_p]),
    "vssr_fire_init": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "vssr_fire_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p,
                               c_int, c_double, c_void_p]),
    "vssr_painn_relax_workspace_bytes": (c_size_t, [c_int, c_int, c_i64]),
    "vssr_painn_relax": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_void_p, c_int, c_int, c_int, c_float, c_float, c_int, c_double, c_i64, c_void_p, c_int,
                                 c_i64, c_int, c_void_p, c_size_t, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "vssr_painn_edge_stats": (c_int, [c_void_p, c_int, c_int, c_i64, c_void_p, c_int, c_void_p, c_void_p, c_int, c_i64,
                                      c_void_p, c_void_p]),
    "vssr_painn_relax_edge_stats": (c_int, [c_void_p, c_int, c_int, c_i64, c_void_p, c_int, c_void_p, c_int, c_i64,
                                            c_void_p, c_void_p]),
    "vssr_classical_smem_bytes": (c_size_t, [c_int, c_int, c_int]),
    "vssr_classical_energy_forces": (c_int, [c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                             c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                             c_void_p, c_void_p]),
    "vssr_classical_relax": (c_int, [c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_void_p, c_int, c_int, c_int, c_int, c_double, c_double, c_void_p,
                                     c_void_p, c_void_p, c_void_p]),
    "vssr_classical_relax_host": (c_int, [c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                          c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_double,
                                          c_double, c_void_p, c_void_p, c_void_p]),
}


class VssrError(RuntimeError):
    pass


def load():
    """Load (once) and return the ctypes handle; raises if the library was not built."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("VSSR_B200_LIB", str(LIB_PATH))
    if not Path(path).exists():
        raise VssrError(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(make -C surface_sampling_b200/csrc). There is no CPU fallback.")
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        kind = {-1: "bad argument", -2: "workspace too small", -3: "unsupported"}.get(rc, f"cudaError {rc}")
        raise VssrError(f"{what} failed: {kind}")
