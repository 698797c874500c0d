"""Batch engines over libvssr_b200.so: many independent chains' slabs resident in HBM.

PyTorch is plumbing only here (device memory, streams); all arithmetic runs in the hand-written
CUDA kernels behind the C ABI (include/vssr_b200.h).  A *batch* is the ragged concatenation of
the chains' structures ("ragged flat" layout, see the header).

Reference mapping:
  PainnEngine.energy_forces  ~ EnsembleNFF.calculate           (mcmc/calculators/calculators.py:484)
  PainnEngine.relax          ~ optimize_slab(optimizer="FIRE") (mcmc/dynamics.py:83-170)
  ClassicalEngine.*          ~ LAMMMPSCalc.run_lammps_energy/_opt (mcmc/calculators/calculators.py:600-640)
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from . import _lib

KCAL_PER_EV = 23.06052
HARTREE_TO_KCAL_MOL = 627.509
HARTREE_TO_EV = HARTREE_TO_KCAL_MOL / KCAL_PER_EV

_PERIODIC = ("H He Li Be B C N O F Ne Na Mg Al Si P S Cl Ar K Ca Sc Ti V Cr Mn Fe Co Ni Cu Zn Ga Ge As Se Br Kr "
             "Rb Sr Y Zr Nb Mo Tc Ru Rh Pd Ag Cd In Sn Sb Te I Xe Cs Ba La Ce Pr Nd Pm Sm Eu Gd Tb Dy Ho Er Tm Yb Lu "
             "Hf Ta W Re Os Ir Pt Au Hg Tl Pb Bi Po At Rn").split()
SYMBOLS = {k + 1: s for k, s in enumerate(_PERIODIC)}
NUMBERS = {v: k for k, v in SYMBOLS.items()}

F, F3, NRBF, NCONV, NEMB, FH = 128, 384, 20, 3, 100, 64

# (name, shape) in packing order — mirror of csrc/painn_layout.h
_LAYER_LAYOUT = [("W1T", (F, F)), ("B1", (F,)), ("W2T", (F, F3)), ("B2", (F3,)), ("WDT", (NRBF, F3)), ("BD", (F3,)),
                 ("UVT", (F, 2 * F)), ("W3T", (2 * F, F)), ("B3", (F,)), ("W4T", (F, F3)), ("B4", (F3,)),
                 ("W1", (F, F)), ("W2", (F3, F)), ("UV", (2 * F, F)), ("W3", (F, 2 * F)), ("W4", (F3, F))]
_READ_LAYOUT = [("W5T", (F, FH)), ("B5", (FH,)), ("W6", (FH,)), ("B6", (4,)), ("W5", (FH, F))]


def painn_weight_floats() -> int:
    n = NEMB * F
    n += NCONV * sum(int(np.prod(s)) for _, s in _LAYER_LAYOUT)
    n += sum(int(np.prod(s)) for _, s in _READ_LAYOUT)
    return 3 * n   # exact + TF32 hi + lo copies


def pack_painn_weights(state: dict) -> np.ndarray:
    """Pack one NFF-PaiNN state dict (keys as in the checkpoints) into the
    flat fp32 layout of csrc/painn_layout.h."""
    g = lambda k: np.asarray(state[k], dtype=np.float32)
    parts = [g("embed_block.atom_embed.weight").reshape(NEMB, F)]
    for l in range(NCONV):
        p = f"message_blocks.{l}.inv_message."
        u = f"update_blocks.{l}."
        W1, b1 = g(p + "inv_dense.layers.0.weight"), g(p + "inv_dense.layers.0.bias")
        W2, b2 = g(p + "inv_dense.layers.1.weight"), g(p + "inv_dense.layers.1.bias")
        Wd, bd = g(p + "dist_embed.block.1.weight"), g(p + "dist_embed.block.1.bias")
        U, V = g(u + "u_mat.weight"), g(u + "v_mat.weight")
        W3, b3 = g(u + "s_dense.0.weight"), g(u + "s_dense.0.bias")
        W4, b4 = g(u + "s_dense.1.weight"), g(u + "s_dense.1.bias")
        t = {
            "W1T": W1.T, "B1": b1, "W2T": W2.T, "B2": b2, "WDT": Wd.T, "BD": bd,
            "UVT": np.concatenate([U.T, V.T], axis=1), "W3T": W3.T, "B3": b3, "W4T": W4.T, "B4": b4,
            "W1": W1, "W2": W2, "UV": np.concatenate([U, V], axis=0), "W3": W3, "W4": W4,
        }
        for name, shape in _LAYER_LAYOUT:
            a = np.ascontiguousarray(t[name], dtype=np.float32)
            assert a.shape == shape, (name, a.shape, shape)
            parts.append(a)
    r = "readout_blocks.0.readoutdict.energy."
    W5, b5, W6, b6 = g(r + "0.weight"), g(r + "0.bias"), g(r + "1.weight"), g(r + "1.bias")
    t = {"W5T": W5.T, "B5": b5, "W6": W6.reshape(FH), "B6": np.array([b6.reshape(-1)[0], 0, 0, 0], np.float32),
         "W5": W5}
    for name, shape in _READ_LAYOUT:
        a = np.ascontiguousarray(t[name], dtype=np.float32)
        assert a.shape == shape, (name, a.shape, shape)
        parts.append(a)
    flat = np.concatenate([x.reshape(-1) for x in parts]).astype(np.float32)
    assert flat.size * 3 == painn_weight_floats()
    # [exact | TF32 part | remainder]: operands of the 3xTF32 tcgen05 GEMM (csrc/gemm_tc.cuh)
    def tf32_rn(x):
        bits = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
        return (((bits + 0x0FFF + ((bits >> 13) & 1)) & 0xFFFFE000) & 0xFFFFFFFF).astype(np.uint32).view(np.float32)

    hi = tf32_rn(flat)
    lo = tf32_rn((flat - hi).astype(np.float32))   # exactly representable: no truncation inside the tensor core
    return np.concatenate([flat, hi, lo])


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _require_cuda():
    if not torch.cuda.is_available():
        raise _lib.VssrError("a CUDA device is required: the engine has no CPU fallback")


@dataclass
class Batch:
    """Ragged batch of structures resident on the device."""
    pos: torch.Tensor        # [A,3] float64
    z: torch.Tensor          # [A] int32 (atomic numbers, or type indices for classical potentials)
    fixed: torch.Tensor      # [A] uint8
    atom_ptr: torch.Tensor   # [B+1] int32
    cell: torch.Tensor       # [B,3,3] float64
    pbc: torch.Tensor        # [B,3] uint8
    n_struct: int
    n_atoms: int
    atom_ptr_host: np.ndarray
    fixed_host: np.ndarray | None = None   # host copy of `fixed` (contract checks only)

    @property
    def cell32(self):
        return self.cell.to(torch.float32)

    @property
    def max_atoms(self) -> int:
        return int(np.diff(self.atom_ptr_host).max()) if self.n_struct else 0

    @staticmethod
    def from_arrays(pos_list, z_list, cell_list, pbc_list, fixed_list=None, device="cuda", pinned=True):
        """Host arrays (one entry per chain) -> one pinned staging buffer each -> device."""
        n = [len(p) for p in pos_list]
        B = len(n)
        ptr = np.zeros(B + 1, dtype=np.int32)
        ptr[1:] = np.cumsum(n)
        A = int(ptr[-1])
        pos = np.concatenate([np.asarray(p, dtype=np.float64).reshape(-1, 3) for p in pos_list]) if A else np.zeros((0, 3))
        z = np.concatenate([np.asarray(q, dtype=np.int32).reshape(-1) for q in z_list]) if A else np.zeros(0, np.int32)
        if fixed_list is None:
            fixed = np.zeros(A, dtype=np.uint8)
        else:
            fixed = np.concatenate([np.asarray(q).astype(np.uint8).reshape(-1) for q in fixed_list])
        cell = np.stack([np.asarray(c, dtype=np.float64).reshape(3, 3) for c in cell_list])
        pbc = np.stack([np.asarray(p).astype(np.uint8).reshape(3) for p in pbc_list])
        return Batch.from_flat(pos, z, fixed, ptr, cell, pbc, device=device, pinned=pinned)

    @staticmethod
    def from_flat(pos, z, fixed, ptr, cell, pbc, device="cuda", pinned=True):
        def up(a, dt):
            t = torch.from_numpy(np.ascontiguousarray(a))
            if pinned and torch.cuda.is_available():
                t = t.pin_memory()
            return t.to(device=device, dtype=dt, non_blocking=True)
        ptr = np.asarray(ptr, dtype=np.int32)
        return Batch(pos=up(pos, torch.float64), z=up(z, torch.int32), fixed=up(fixed, torch.uint8),
                     atom_ptr=up(ptr, torch.int32), cell=up(cell, torch.float64), pbc=up(pbc, torch.uint8),
                     n_struct=len(ptr) - 1, n_atoms=int(ptr[-1]), atom_ptr_host=ptr.copy(),
                     fixed_host=np.asarray(fixed).astype(bool))

    def h2d_bytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in (self.pos, self.z, self.fixed, self.atom_ptr, self.cell, self.pbc))

    def split_host(self, flat: np.ndarray):
        p = self.atom_ptr_host
        return [flat[p[b]:p[b + 1]] for b in range(self.n_struct)]


class _Workspace:
    def __init__(self):
        self.buf = None

    def get(self, nbytes: int, device) -> torch.Tensor:
        # (an engine lives on one device; comparing torch.device('cuda') with cuda:0 is always unequal)
        if self.buf is None or self.buf.numel() < nbytes:
            # 1.5x head room: MC chains gain adsorbates step by step, and re-allocating a GB-sized workspace
            # (cudaFree synchronises the device, cudaMalloc takes tens of ms) must stay a rare event
            self.buf = None
            self.buf = torch.empty(int(nbytes * 1.5) + 256, dtype=torch.uint8, device=device)
        return self.buf


def neighbor_list(batch: Batch, cutoff: float, e_cap: int | None = None):
    """Directed periodic neighbour list (vssr_nbr_build). Returns (rowptr[A+1], col[E], shift[E,4]) on
    the device, exactly sized (re-runs once if the first capacity guess was too small)."""
    _require_cuda()
    lib = _lib.load()
    dev = batch.pos.device
    A, B = batch.n_atoms, batch.n_struct
    pos32 = batch.pos.to(torch.float32)
    cell32 = batch.cell32.contiguous()
    cap = int(e_cap) if e_cap else max(A * 96, 1024)
    while True:
        deg = torch.empty(max(A, 1), dtype=torch.int32, device=dev)
        rowptr = torch.empty(A + 1, dtype=torch.int32, device=dev)
        col = torch.empty(cap, dtype=torch.int32, device=dev)
        shift = torch.empty((cap, 4), dtype=torch.int8, device=dev)
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        _lib.check(lib.vssr_nbr_build(_ptr(pos32), _ptr(batch.atom_ptr), _ptr(cell32), _ptr(batch.pbc), B, A,
                                      float(cutoff), _ptr(deg), _ptr(rowptr), _ptr(col), _ptr(shift), cap,
                                      _ptr(status), _stream()), "vssr_nbr_build")
        n_edges = int(rowptr[-1].item())
        if n_edges <= cap:
            return rowptr, col[:n_edges], shift[:n_edges]
        cap = n_edges


class PainnEngine:
    """3-model (or M-model) PaiNN ensemble on the GPU."""

    def __init__(self, states: list[dict], offset_data: dict | None = None, cutoff: float = 5.0, skin: float = 1.0,
                 device: str = "cuda", edges_per_atom: int = 96):
        _require_cuda()
        self.lib = _lib.load()
        assert int(self.lib.vssr_painn_weight_floats()) == painn_weight_floats(), "weight layout mismatch"
        self.device = torch.device(device)
        self.n_models = len(states)
        w = np.stack([pack_painn_weights(s) for s in states])
        self.weights = torch.from_numpy(w).to(self.device)
        self.offset_data = offset_data
        self.cutoff, self.skin = float(cutoff), float(skin)
        self.edges_per_atom = edges_per_atom
        self._ws = _Workspace()
        self._ws_relax = _Workspace()
        self._fc = None          # (blob tensor, n0, e_cap0, nslots) of the frozen-pair filter memo

    def set_framework(self, pos0, cell, pbc, fixed0, constrained_forces: bool = False, pair_kernels: bool = True) -> int:
        """Build the radial-filter memo for a frozen framework shared by every structure (its atoms must
        be the FIRST n0 atoms of each structure, as in VSSR-MC where adsorbates are appended).  Edges
        whose distance is bitwise equal to a framework edge between two frozen atoms reuse the memoised
        filter rows instead of re-evaluating 2x60 FMAs per feature; anything else is computed as usual,
        so correctness never depends on this call.  Returns the number of memoised pairs.

        constrained_forces=True (VSSR_FC_CONSTRAINED_GRAD) additionally declares that the frozen atoms
        carry FixAtoms, as in every VSSR-MC relaxation (mcmc/dynamics.py:25-141 -> ASE zeroes their
        forces): dE/dx of those atoms is then not computed at all (their force rows come back as 0),
        which removes the whole position-gradient branch of every memoised edge.  Energies and the
        forces of all other atoms are unchanged.  pair_kernels=False (VSSR_FC_NO_PAIR) keeps the
        one-structure-per-CTA memo kernels (same bits; for tests)."""
        lib, dev = self.lib, self.device
        pos0 = np.ascontiguousarray(pos0, dtype=np.float32)
        n0 = len(pos0)
        b = Batch.from_arrays([pos0], [np.zeros(n0, np.int32)], [cell], [pbc])
        rowptr, col, _ = neighbor_list(b, self.cutoff + self.skin)
        e_cap0 = int(col.numel()) + 8
        nbytes = int(lib.vssr_painn_filter_cache_bytes(self.n_models, n0, e_cap0))
        blob = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        ws = torch.empty(int(lib.vssr_painn_filter_cache_workspace_bytes(n0, e_cap0)), dtype=torch.uint8, device=dev)
        d_pos = torch.from_numpy(pos0).to(dev)
        d_cell = torch.from_numpy(np.ascontiguousarray(cell, dtype=np.float32)).to(dev)
        d_pbc = torch.from_numpy(np.ascontiguousarray(pbc).astype(np.uint8)).to(dev)
        d_fix = torch.from_numpy(np.ascontiguousarray(fixed0).astype(np.uint8)).to(dev)
        import ctypes
        nslots = ctypes.c_int32(0)
        _lib.check(lib.vssr_painn_filter_cache_build(_ptr(self.weights), self.n_models, _ptr(d_pos), _ptr(d_cell),
                                                     _ptr(d_pbc), _ptr(d_fix), n0, self.cutoff, self.skin, e_cap0,
                                                     _ptr(blob), blob.numel(), _ptr(ws), ws.numel(),
                                                     ctypes.addressof(nslots), _stream()),
                   "vssr_painn_filter_cache_build")
        self._fc = (blob, n0, e_cap0, int(nslots.value), (1 if constrained_forces else 0) | (0 if pair_kernels else 2),
                    np.ascontiguousarray(fixed0).astype(bool))
        return int(nslots.value)

    def clear_framework(self):
        self._fc = None

    def _result_buffers(self, B: int, A: int, into: dict | None):
        """Output tensors of one relax() call: caller-owned when `into` is given (keys out[B,8] f64, forces[A,3] f32,
        forces_std[A,3] f32, status[1] i32 -- validated here), else freshly allocated per call (torch's caching
        allocator: no cudaMalloc after the first MC steps).  A result is never aliased by a later call."""
        dev = self.device
        if into is None:
            return (torch.empty((B, 8), dtype=torch.float64, device=dev), torch.empty((A, 3), dtype=torch.float32, device=dev),
                    torch.empty((A, 3), dtype=torch.float32, device=dev), torch.empty(1, dtype=torch.int32, device=dev))
        want = {"out": ((B, 8), torch.float64), "forces": ((A, 3), torch.float32), "forces_std": ((A, 3), torch.float32),
                "status": ((1,), torch.int32)}
        for k, (shape, dt) in want.items():
            t = into.get(k)
            if t is None or tuple(t.shape) != shape or t.dtype != dt or not t.is_contiguous() or t.device.type != dev.type:
                raise _lib.VssrError(f"relax(into=...): `{k}` must be a contiguous {dt} tensor of shape {shape} on {dev}")
        return into["out"], into["forces"], into["forces_std"], into["status"]

    def _fc_args(self, constrained: bool = True):
        if self._fc is None:
            return None, 0, 0, 0
        flags = self._fc[4] if constrained else (self._fc[4] & ~1)
        return self._fc[0].data_ptr(), self._fc[1], self._fc[2], flags

    # -- H5 stoichiometric offset (per structure, host, exact fp64) ---------------------------
    def offsets_ev(self, z_host: np.ndarray, atom_ptr: np.ndarray) -> np.ndarray | None:
        if self.offset_data is None:
            return None
        sto = self.offset_data["stoidict"]
        z_host = np.asarray(z_host)
        # exact integer element counts per structure -> the result is independent of what else is
        # in the batch (bitwise batch invariance), unlike a float cumsum over atoms
        tot = np.full(len(atom_ptr) - 1, float(sto["offset"]))
        for zz in sorted(SYMBOLS):
            sym = SYMBOLS[zz]
            if sym not in sto:
                continue
            csum = np.concatenate([[0], np.cumsum(z_host == zz, dtype=np.int64)])
            cnt = csum[atom_ptr[1:]] - csum[atom_ptr[:-1]]
            tot = tot + cnt * float(sto[sym])
        return tot * HARTREE_TO_KCAL_MOL / KCAL_PER_EV

    def energy_forces(self, batch: Batch, z_host: np.ndarray | None = None, want_embedding=False, nbrs=None,
                      constrained_forces: bool = False):
        """One ensemble evaluation of every structure in the batch.  Forces are the raw forces on every
        atom (what EnsembleNFF.calculate returns) unless constrained_forces=True AND the framework was
        registered with constrained_forces=True: then the force rows of its frozen atoms are zero."""
        lib, dev = self.lib, self.device
        A, B, M = batch.n_atoms, batch.n_struct, self.n_models
        if nbrs is None:
            nbrs = neighbor_list(batch, self.cutoff + self.skin)
        rowptr, col, shift = nbrs
        e_cap = max(int(col.numel()), 1)
        if col.numel() == 0:
            col = torch.zeros(1, dtype=torch.int32, device=dev)
            shift = torch.zeros((1, 4), dtype=torch.int8, device=dev)
        pos32 = batch.pos.to(torch.float32)
        cell32 = batch.cell32.contiguous()
        nbytes = int(lib.vssr_painn_workspace_bytes(M, A, e_cap))
        ws = self._ws.get(nbytes, dev)
        energy = torch.empty((M, B), dtype=torch.float64, device=dev)
        grad = torch.empty((M, A, 3), dtype=torch.float32, device=dev)
        emb = torch.empty((M, A, F), dtype=torch.float32, device=dev) if want_embedding else None
        _lib.check(lib.vssr_painn_energy_grad(_ptr(self.weights), M, _ptr(pos32), _ptr(batch.z), _ptr(batch.atom_ptr),
                                              _ptr(cell32), B, A, batch.max_atoms, _ptr(rowptr), _ptr(col), _ptr(shift), e_cap,
                                              self.cutoff, *self._fc_args(constrained_forces), _ptr(ws), ws.numel(), _ptr(energy), _ptr(grad), _ptr(emb),
                                              _stream()), "vssr_painn_energy_grad")
        off = None
        if self.offset_data is not None:
            zh = z_host if z_host is not None else batch.z.cpu().numpy()
            off = torch.from_numpy(self.offsets_ev(zh, batch.atom_ptr_host)).to(dev)
        e_mean = torch.empty(B, dtype=torch.float64, device=dev)
        e_std = torch.empty(B, dtype=torch.float64, device=dev)
        f_mean = torch.empty((A, 3), dtype=torch.float32, device=dev)
        f_std = torch.empty((A, 3), dtype=torch.float32, device=dev)
        _lib.check(lib.vssr_ensemble_stats(_ptr(energy), _ptr(grad), _ptr(off), _ptr(batch.atom_ptr), M, B, A,
                                           _ptr(e_mean), _ptr(e_std), _ptr(f_mean), _ptr(f_std), _stream()),
                   "vssr_ensemble_stats")
        return {"energy": e_mean, "energy_std": e_std, "forces": f_mean, "forces_std": f_std,
                "energy_kcal_per_model": energy, "grad_kcal_per_model": grad, "embedding": emb}

    def relax(self, batch: Batch, relax_steps: int = 20, fmax: float = 0.01, z_host: np.ndarray | None = None,
              want_std: bool = True, e_cap: int | None = None, check: bool = False, into: dict | None = None):
        """optimize_slab(optimizer='FIRE') for every structure, no host round trip.  batch.pos is
        updated in place.  Returns dict(out[B,8], forces, forces_std, status).

        The call is asynchronous; `status` (device int) carries VSSR_STATUS_EDGE_OVERFLOW if the neighbour list
        needed more than e_cap edges (default: edges_per_atom per atom).  check=True synchronises, and on overflow
        restores the positions and repeats the relaxation with a doubled capacity."""
        if check:
            pos0 = batch.pos.clone()
            cap_try = int(e_cap) if e_cap else batch.n_atoms * self.edges_per_atom
            while True:
                r = self.relax(batch, relax_steps, fmax, z_host, want_std, cap_try, check=False, into=into)
                st = int(r["status"].item())
                if not (st & 1):
                    return r
                batch.pos.copy_(pos0)
                cap_try *= 2
        lib, dev = self.lib, self.device
        A, B, M = batch.n_atoms, batch.n_struct, self.n_models
        if self._fc is not None and (self._fc[4] & 1) and batch.fixed_host is not None:
            # VSSR_FC_CONSTRAINED_GRAD contract: the framework's frozen atoms are FixAtoms in every structure
            frozen = np.flatnonzero(self._fc[5])
            idx = (batch.atom_ptr_host[:-1, None] + frozen[None, :]).reshape(-1)
            if (np.diff(batch.atom_ptr_host) < self._fc[1]).any() or not batch.fixed_host[idx].all():
                raise _lib.VssrError("set_framework(constrained_forces=True): a structure does not hold the framework's "
                                     "frozen atoms fixed; use constrained_forces=False")
        cap = int(e_cap) if e_cap else A * self.edges_per_atom
        nbytes = int(lib.vssr_painn_relax_workspace_bytes(M, A, cap))
        ws = self._ws_relax.get(nbytes, dev)
        off = None
        if self.offset_data is not None:
            zh = z_host if z_host is not None else batch.z.cpu().numpy()
            off = torch.from_numpy(self.offsets_ev(zh, batch.atom_ptr_host)).to(dev, non_blocking=True)
        out, forces, fstd_buf, status = self._result_buffers(B, A, into)
        fstd = fstd_buf if want_std else None
        status.zero_()
        cell32 = batch.cell32.contiguous()
        _lib.check(lib.vssr_painn_relax(_ptr(self.weights), M, _ptr(batch.pos), _ptr(batch.z), _ptr(batch.fixed),
                                        _ptr(batch.atom_ptr), _ptr(cell32), _ptr(batch.pbc), _ptr(off), B, A,
                                        batch.max_atoms, self.cutoff, self.skin, int(relax_steps), float(fmax), cap,
                                        *self._fc_args(), _ptr(ws), ws.numel(), _ptr(out), _ptr(forces), _ptr(fstd), _ptr(status), _stream()),
                   "vssr_painn_relax")
        self._last_relax = (A, cap, M, ws)
        return {"out": out, "forces": forces, "forces_std": fstd, "status": status}

    def last_relax_edge_stats(self, batch: Batch) -> dict:
        """Work done by the LAST evaluation of the most recent relax() on `batch` (same atom count): numbers of direct
        and memoised edges etc. (vssr_painn_relax_edge_stats).  For bench.py's executed-work roofline; synchronises."""
        A, cap, M, ws = self._last_relax
        assert A == batch.n_atoms, "edge stats refer to the most recent relax() call"
        out = torch.zeros(5, dtype=torch.int64, device=self.device)
        fcp, n0, ecap0, _ = self._fc_args()
        _lib.check(self.lib.vssr_painn_relax_edge_stats(_ptr(ws), M, A, cap, _ptr(batch.atom_ptr), batch.n_struct, fcp, n0,
                                                        ecap0, _ptr(out), _stream()), "vssr_painn_relax_edge_stats")
        d, m, df, c, e6 = (int(x) for x in out.cpu().tolist())
        return {"direct_edges": d, "memo_edges": m, "direct_edges_frozen_receiver": df, "canonical_structures": c,
                "skin_list_edges": e6, "atoms": A, "structures": batch.n_struct}


RELAX_COLS = ("energy", "energy_std", "raw_energy", "max_abs_force", "nsteps", "converged", "energy_oob", "n_evals")

POT_TERSOFF, POT_SW, POT_EAM = 0, 1, 2


def tersoff_param_table(pot_json: dict, elements: list[str]) -> np.ndarray:
    """[ntypes^3,14] table (LAMMPS elem3param order) from the parsed GaN.tersoff fixture."""
    ne = len(elements)
    tab = np.zeros((ne, ne, ne, 14))
    seen = np.zeros((ne, ne, ne), bool)
    for e in pot_json["entries"]:
        if all(x in elements for x in e["elements"]):
            a, b, c = (elements.index(x) for x in e["elements"])
            tab[a, b, c] = e["params"]
            seen[a, b, c] = True
    if not seen.all():
        raise ValueError("missing Tersoff entries for " + str(elements))
    return tab.reshape(-1, 14)


def sw_param_table(eps=2.1683, sigma=2.0951, a=1.80, lam=21.0, gamma=1.20, costheta0=-1.0 / 3.0, A=7.049556277,
                   B=0.6022245584, p=4.0, q=0.0) -> np.ndarray:
    """Single-element SW table [1,10] (LAMMPS `pair_style sw` order); SW-1985 Si by default."""
    return np.array([[eps, sigma, a, lam, gamma, costheta0, A, B, p, q]], dtype=np.float64)


def _eam_spline(f: np.ndarray, delta: float) -> np.ndarray:
    """The 7-coefficient cubic-spline table LAMMPS builds for every EAM function (PairEAM::interpolate; SURVEY.md
    App. A.4): rows 1..n, row 0 unused.  Columns 6..3 = value and cubic coefficients in the reduced coordinate,
    columns 2..0 = the derivative polynomial (already divided by the grid spacing)."""
    n = len(f)
    c = np.zeros((n + 1, 7))
    c[1:, 6] = f
    c[1, 5] = c[2, 6] - c[1, 6]
    c[2, 5] = 0.5 * (c[3, 6] - c[1, 6])
    c[n - 1, 5] = 0.5 * (c[n, 6] - c[n - 2, 6])
    c[n, 5] = c[n, 6] - c[n - 1, 6]
    k = np.arange(3, n - 1)
    c[k, 5] = ((c[k - 2, 6] - c[k + 2, 6]) + 8.0 * (c[k + 1, 6] - c[k - 1, 6])) / 12.0
    k = np.arange(1, n)
    rise = c[k + 1, 6] - c[k, 6]
    c[k, 4] = 3.0 * rise - 2.0 * c[k, 5] - c[k + 1, 5]
    c[k, 3] = c[k, 5] + c[k + 1, 5] - 2.0 * rise
    c[1:, 2] = c[1:, 5] / delta
    c[1:, 1] = 2.0 * c[1:, 4] / delta
    c[1:, 0] = 3.0 * c[1:, 3] / delta
    return c


def eam_param_block(funcfl: dict) -> np.ndarray:
    """Parameter block of VSSR_POT_EAM (include/vssr_b200.h) from a parsed single-element funcfl file
    (loaders.load_eam_funcfl): header + spline tables of F(rho), rho(r) and z2r(r) = 27.2*0.529*Z(r)^2 = r*phi(r)."""
    nrho, drho, nr, dr, rc = int(funcfl["nrho"]), float(funcfl["drho"]), int(funcfl["nr"]), float(funcfl["dr"]), float(funcfl["rc"])
    zr = np.asarray(funcfl["zr"], dtype=np.float64)
    tabs = [_eam_spline(np.asarray(funcfl["frho"], dtype=np.float64), drho),
            _eam_spline(np.asarray(funcfl["rhor"], dtype=np.float64), dr),
            _eam_spline(27.2 * 0.529 * zr * zr, dr)]
    return np.concatenate([np.array([nrho, drho, nr, dr, rc, 0.0, 0.0, 0.0])] + [t.reshape(-1) for t in tabs])


class ClassicalEngine:
    """Tersoff / Stillinger-Weber / EAM energy, forces and FIRE relaxation, one CTA per chain."""

    def __init__(self, kind: int, params: np.ndarray, ntypes: int, n_max: int = 128, max_nbr: int = 32,
                 skin: float = 0.5, device: str = "cuda"):
        _require_cuda()
        self.lib = _lib.load()
        self.kind, self.ntypes, self.n_max, self.max_nbr, self.skin = kind, ntypes, n_max, max_nbr, skin
        self.device = torch.device(device)
        self.params = torch.from_numpy(np.ascontiguousarray(params, dtype=np.float64)).to(self.device)
        smem = int(self.lib.vssr_classical_smem_bytes(kind, n_max, max_nbr))
        if smem > 227 * 1024:
            raise ValueError(f"n_max={n_max}, max_nbr={max_nbr} needs {smem} B shared memory (> 227 KB)")

    def _check_status(self, status):
        s = int(status.item())
        if s & 2:
            raise _lib.VssrError("classical kernel: neighbour slots overflowed (raise max_nbr; Tersoff / SW also keep at most "
                                 "8 * n_max atom pairs inside the potential cutoff: raise n_max)")
        if s & 4:
            raise _lib.VssrError("classical kernel: structure larger than n_max")

    def energy_forces(self, batch: Batch, check=True):
        dev = self.device
        A, B = batch.n_atoms, batch.n_struct
        energy = torch.empty(B, dtype=torch.float64, device=dev)
        forces = torch.empty((A, 3), dtype=torch.float64, device=dev)
        eat = torch.empty(A, dtype=torch.float64, device=dev)
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        _lib.check(self.lib.vssr_classical_energy_forces(self.kind, _ptr(self.params), self.ntypes, _ptr(batch.pos),
                                                         _ptr(batch.z), _ptr(batch.atom_ptr), _ptr(batch.cell),
                                                         _ptr(batch.pbc), B, self.n_max, self.max_nbr, _ptr(energy),
                                                         _ptr(forces), _ptr(eat), _ptr(status), _stream()),
                   "vssr_classical_energy_forces")
        if check:
            self._check_status(status)
        return {"energy": energy, "forces": forces, "per_atom_energies": eat, "status": status}

    def relax(self, batch: Batch, relax_steps: int = 100, fmax: float = 0.01, check=True):
        dev = self.device
        A, B = batch.n_atoms, batch.n_struct
        out = torch.empty((B, 8), dtype=torch.float64, device=dev)
        forces = torch.empty((A, 3), dtype=torch.float64, device=dev)
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        _lib.check(self.lib.vssr_classical_relax(self.kind, _ptr(self.params), self.ntypes, _ptr(batch.pos),
                                                 _ptr(batch.z), _ptr(batch.fixed), _ptr(batch.atom_ptr),
                                                 _ptr(batch.cell), _ptr(batch.pbc), B, self.n_max, self.max_nbr,
                                                 int(relax_steps), float(fmax), float(self.skin), _ptr(out),
                                                 _ptr(forces), _ptr(status), _stream()), "vssr_classical_relax")
        if check:
            self._check_status(status)
        return {"out": out, "forces": forces, "status": status}


def system_reduce(per_atom: torch.Tensor, batch: Batch) -> torch.Tensor:
    """get_system_val (mcmc/uncertainty/prediction.py:181-223): [B,6] = sum,max,min,mean,mean_sq,rms."""
    lib = _lib.load()
    out = torch.empty((batch.n_struct, 6), dtype=torch.float32, device=per_atom.device)
    _lib.check(lib.vssr_system_reduce(_ptr(per_atom.contiguous()), _ptr(batch.atom_ptr), batch.n_struct, _ptr(out),
                                      _stream()), "vssr_system_reduce")
    return out


def atom_norm(vec: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    out = torch.empty(vec.shape[0], dtype=torch.float32, device=vec.device)
    _lib.check(lib.vssr_atom_norm(_ptr(vec.contiguous()), vec.shape[0], _ptr(out), _stream()), "vssr_atom_norm")
    return out
