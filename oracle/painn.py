"""ORACLE (test infrastructure only — never imported by the product path).

CPU restatement (PyTorch, fp32 or fp64, autograd forces) of the PaiNN model the reference
evaluates through NFF: ``nff.nn.models.painn.Painn`` as pickled in
``tutorials/data/SrTiO3_001/nff/model0{1,2,3}/best_model`` and called from
``EnsembleNFF.calculate`` <- ``mcmc/calculators/calculators.py:484``.
NFF is an un-vendored dependency (``nff @ surface-sampling-0.3.0``, pyproject.toml:17); the
algorithm below restates its published PaiNN and is PINNED against the
golden numbers printed in the reference notebooks (tests/test_oracle_golden.py):
  tutorials/SrTiO3_001.ipynb:241           -467.521881 eV, fmax 0.204407 eV/A
  tests/test_SrTiO3_terms.ipynb:201,208,212 -570.127991 / -467.525604 / -518.694092 eV

Weights are addressed with the checkpoint's own state-dict keys.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from .nbrlist import neighbor_list

EV_TO_KCAL_MOL = 23.06052   # nff.utils.constants
HARTREE_TO_KCAL_MOL = 627.509
HARTREE_TO_EV = HARTREE_TO_KCAL_MOL / EV_TO_KCAL_MOL

FEAT = 128
N_RBF = 20
N_CONV = 3
CUTOFF = 5.0
SKIN = 1.0
VEX_SIGMA = 1.5
VEX_POWER = 12

_SYMBOL = {1: "H", 8: "O", 22: "Ti", 38: "Sr", 57: "La", 25: "Mn"}


def swish(x):
    return x * torch.sigmoid(x)


def init_random_weights(seed: int) -> dict[str, np.ndarray]:
    """Random-init PaiNN with the checkpoint's shapes (BASELINE.json: checkpoints 'unavailable
    offline' -> throughput runs use xavier-uniform weights, zero biases, N(0,1) embedding with
    row 0 zero)."""
    g = torch.Generator().manual_seed(seed)

    def xavier(o, i):
        a = math.sqrt(6.0 / (i + o))
        return ((torch.rand(o, i, generator=g) * 2 - 1) * a).numpy().astype(np.float32)

    sd = {}
    emb = torch.randn(100, FEAT, generator=g).numpy().astype(np.float32)
    emb[0] = 0
    sd["embed_block.atom_embed.weight"] = emb
    for l in range(N_CONV):
        p = f"message_blocks.{l}.inv_message."
        sd[p + "inv_dense.layers.0.weight"] = xavier(FEAT, FEAT)
        sd[p + "inv_dense.layers.0.bias"] = np.zeros(FEAT, np.float32)
        sd[p + "inv_dense.layers.1.weight"] = xavier(3 * FEAT, FEAT)
        sd[p + "inv_dense.layers.1.bias"] = np.zeros(3 * FEAT, np.float32)
        sd[p + "dist_embed.block.1.weight"] = xavier(3 * FEAT, N_RBF)
        sd[p + "dist_embed.block.1.bias"] = np.zeros(3 * FEAT, np.float32)
        u = f"update_blocks.{l}."
        sd[u + "u_mat.weight"] = xavier(FEAT, FEAT)
        sd[u + "v_mat.weight"] = xavier(FEAT, FEAT)
        sd[u + "s_dense.0.weight"] = xavier(FEAT, 2 * FEAT)
        sd[u + "s_dense.0.bias"] = np.zeros(FEAT, np.float32)
        sd[u + "s_dense.1.weight"] = xavier(3 * FEAT, FEAT)
        sd[u + "s_dense.1.bias"] = np.zeros(3 * FEAT, np.float32)
    r = "readout_blocks.0.readoutdict.energy."
    sd[r + "0.weight"] = xavier(FEAT // 2, FEAT)
    sd[r + "0.bias"] = np.zeros(FEAT // 2, np.float32)
    sd[r + "1.weight"] = xavier(1, FEAT // 2)
    sd[r + "1.bias"] = np.zeros(1, np.float32)
    return sd


def load_golden_weights(path) -> list[dict[str, np.ndarray]]:
    z = np.load(path)
    models = []
    for m in ("model01", "model02", "model03"):
        models.append({k[len(m) + 1:]: z[k] for k in z.files if k.startswith(m + "/")})
    return models


class PainnOracle:
    """One PaiNN model; ``energy_and_grad`` returns kcal/mol like the NFF module does."""

    def __init__(self, state: dict[str, np.ndarray], dtype=torch.float32, cutoff: float = CUTOFF, device="cpu"):
        # device="cuda": the same eager torch code on the GPU -- bench.py's "reference-equivalent GPU path" baseline
        # (BASELINE.md section 2, config 4); everything else in the repo uses the CPU default
        self.dtype = dtype
        self.cutoff = cutoff
        self.device = torch.device(device)
        self.w = {k: torch.tensor(np.asarray(v), dtype=dtype, device=self.device) for k, v in state.items()}

    # -- pieces -------------------------------------------------------------------------
    def _edge_geometry(self, xyz, nbr_i, nbr_j, offsets):
        # nff get_rij: r_ij = xyz[j] - xyz[i] + offsets ; keep dist <= cutoff
        r_ij = xyz[nbr_j] - xyz[nbr_i] + offsets
        dist_plain = (r_ij.detach() ** 2).sum(-1) ** 0.5
        keep = dist_plain <= self.cutoff
        return r_ij[keep], nbr_i[keep], nbr_j[keep]

    def forward_energy(self, xyz, z, nbr_i, nbr_j, offsets, return_features=False):
        w = self.w
        r_ij, ni, nj = self._edge_geometry(xyz, nbr_i, nbr_j, offsets)
        n_atoms = xyz.shape[0]
        # preprocess_r: dist with 1e-10 inside the sum, unit = r/dist
        dist = ((r_ij ** 2 + 1e-10).sum(-1)) ** 0.5
        unit = r_ij / dist.reshape(-1, 1)
        # PainnRadialBasis: sin(n pi d / cutoff) / d, zero for d >= cutoff
        n = torch.arange(1, N_RBF + 1, dtype=self.dtype, device=xyz.device)
        shape_d = dist.unsqueeze(-1)
        coef = n * math.pi / self.cutoff
        denom = torch.where(shape_d == 0, torch.ones_like(shape_d), shape_d)
        rbf = torch.where(shape_d >= self.cutoff, torch.zeros_like(shape_d * coef),
                          torch.sin(coef * shape_d) / denom)
        # CosineEnvelope
        env = 0.5 * (torch.cos(math.pi * dist / self.cutoff) + 1)
        env = torch.where(dist >= self.cutoff, torch.zeros_like(env), env)

        s = w["embed_block.atom_embed.weight"][z]            # [N,F]
        v = torch.zeros(n_atoms, FEAT, 3, dtype=self.dtype, device=xyz.device)  # [N,F,3]
        for l in range(N_CONV):
            p = f"message_blocks.{l}.inv_message."
            h = swish(s @ w[p + "inv_dense.layers.0.weight"].T + w[p + "inv_dense.layers.0.bias"])
            phi = h @ w[p + "inv_dense.layers.1.weight"].T + w[p + "inv_dense.layers.1.bias"]
            wd = rbf @ w[p + "dist_embed.block.1.weight"].T + w[p + "dist_embed.block.1.bias"]
            wd = wd * env.reshape(-1, 1)
            inv_out = (phi[nj] * wd).reshape(-1, 3, FEAT)
            split_0 = inv_out[:, 0, :].unsqueeze(-1)
            split_1 = inv_out[:, 1, :]
            split_2 = inv_out[:, 2, :].unsqueeze(-1)
            unit_add = split_2 * unit.unsqueeze(1)
            delta_v_ij = unit_add + split_0 * v[nj]
            delta_s_ij = split_1
            dv = torch.zeros_like(v).index_add_(0, ni, delta_v_ij)
            ds = torch.zeros_like(s).index_add_(0, ni, delta_s_ij)
            s = s + ds
            v = v + dv
            u = f"update_blocks.{l}."
            # u_mat/v_mat act on the feature dim: v is [N,F,3] -> transpose to [N,3,F]
            v_t = v.transpose(1, 2)
            u_v = (v_t @ w[u + "u_mat.weight"].T).transpose(1, 2)
            v_v = (v_t @ w[u + "v_mat.weight"].T).transpose(1, 2)
            v_v_norm = ((v_v ** 2 + 1e-15).sum(-1)) ** 0.5
            s_stack = torch.cat([s, v_v_norm], dim=-1)
            a = swish(s_stack @ w[u + "s_dense.0.weight"].T + w[u + "s_dense.0.bias"])
            a = a @ w[u + "s_dense.1.weight"].T + w[u + "s_dense.1.bias"]
            split = a.reshape(n_atoms, 3, FEAT)
            a_vv = split[:, 0, :].unsqueeze(-1)
            dv_u = u_v * a_vv
            a_sv = split[:, 1, :]
            a_ss = split[:, 2, :]
            inner = (u_v * v_v).sum(-1)
            ds_u = inner * a_sv + a_ss
            s = s + ds_u
            v = v + dv_u
        r = "readout_blocks.0.readoutdict.energy."
        e_atom = swish(s @ w[r + "0.weight"].T + w[r + "0.bias"]) @ w[r + "1.weight"].T + w[r + "1.bias"]
        e_atom = e_atom.reshape(-1)
        # excluded volume on the cutoff-filtered directed list (with the periodic offsets)
        vex = (VEX_SIGMA / ((r_ij ** 2).sum(1).sqrt())) ** VEX_POWER
        e_atom = e_atom + torch.zeros_like(e_atom).index_add_(0, ni, vex)
        if return_features:
            return e_atom.sum(), s, e_atom
        return e_atom.sum()

    def energy_and_grad(self, pos, numbers, nbr_i, nbr_j, offsets):
        dev = self.device
        xyz = torch.tensor(np.asarray(pos), dtype=self.dtype, device=dev, requires_grad=True)
        z = torch.as_tensor(np.asarray(numbers), dtype=torch.long).to(dev)
        off = torch.tensor(np.asarray(offsets), dtype=self.dtype, device=dev)
        ni = torch.as_tensor(np.asarray(nbr_i), dtype=torch.long).to(dev)
        nj = torch.as_tensor(np.asarray(nbr_j), dtype=torch.long).to(dev)
        e = self.forward_energy(xyz, z, ni, nj, off)
        (g,) = torch.autograd.grad(e, xyz)
        return e.detach().cpu(), g.detach().cpu()


def stoich_offset_kcal(numbers, stoidict: dict) -> float:
    """EnsembleNFF.offset_energy: (sum n_el*stoidict[el] + stoidict['offset']) Ha -> kcal/mol."""
    tot = stoidict["offset"]
    for zz in np.asarray(numbers):
        tot += stoidict[_SYMBOL[int(zz)]]
    return tot * HARTREE_TO_KCAL_MOL


class EnsembleOracle:
    """H5: EnsembleNFF.calculate restated — per-model kcal/mol -> eV, stoichiometric offset,
    mean and population std over models (SURVEY.md App. A.2 'Units/offset')."""

    def __init__(self, states, offset_data: dict | None, dtype=torch.float32, cutoff=CUTOFF,
                 skin=SKIN, device="cpu"):
        self.models = [PainnOracle(s, dtype=dtype, cutoff=cutoff, device=device) for s in states]
        self.offset_data = offset_data
        self.dtype = dtype
        self.cutoff = cutoff
        self.skin = skin
        self.np_dtype = np.float32 if dtype == torch.float32 else np.float64

    def build_nbrs(self, pos, cell, pbc):
        i, j, S = neighbor_list(pos, cell, pbc, self.cutoff + self.skin)
        offsets = (S.astype(np.float64) @ np.asarray(cell, dtype=np.float64)).astype(self.np_dtype)
        return i, j, S, offsets

    def calculate(self, pos, numbers, cell, pbc, nbrs=None) -> dict:
        if nbrs is None:
            nbrs = self.build_nbrs(pos, cell, pbc)
        i, j, _, offsets = nbrs
        es, gs = [], []
        for m in self.models:
            e, g = m.energy_and_grad(np.asarray(pos, dtype=self.np_dtype), numbers, i, j, offsets)
            es.append(e.numpy().astype(self.np_dtype) * self.np_dtype(1 / EV_TO_KCAL_MOL))
            gs.append(g.numpy().astype(self.np_dtype) * self.np_dtype(1 / EV_TO_KCAL_MOL))
        es = np.stack(es).reshape(len(self.models), 1)
        gs = np.stack(gs)
        if self.offset_data is not None:
            off_ev = stoich_offset_kcal(numbers, self.offset_data["stoidict"]) / EV_TO_KCAL_MOL
            es = (es + self.np_dtype(off_ev)).astype(self.np_dtype)
        return {
            "energy": es.mean(0).reshape(-1),
            "energy_std": es.std(0).reshape(-1),
            "forces": -gs.mean(0).reshape(-1, 3),
            "forces_std": gs.std(0).reshape(-1, 3),
            "energies_per_model": es.reshape(-1),
            "grads_per_model": gs,
        }


def surface_energy(energy: float, numbers, offset_data: dict, chem_pots: dict,
                   offset_units: str = "atomic") -> float:
    """H6: EnsembleNFFSurface.get_surface_energy (mcmc/calculators/calculators.py:379-446)."""
    from collections import Counter

    cnt = Counter(_SYMBOL[int(z)] for z in np.asarray(numbers))
    bulk = offset_data["bulk_energies"]
    stoics = offset_data["stoics"]
    ref_formula = offset_data["ref_formula"]
    ref_el = offset_data["ref_element"]
    bulk_ref = cnt[ref_el] * bulk[ref_formula]
    for el in cnt:
        if el != ref_el:
            bulk_ref += (cnt[el] - stoics[el] / stoics[ref_el] * cnt[ref_el]) * bulk[el]
    e = float(energy)
    e -= bulk_ref * HARTREE_TO_EV if offset_units == "atomic" else bulk_ref
    pot = 0.0
    for el in cnt:
        if el != ref_el:
            pot += (cnt[el] - stoics[el] / stoics[ref_el] * cnt[ref_el]) * chem_pots[el]
    return e - pot
