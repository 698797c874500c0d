"""ORACLE-side prototype (test infrastructure only).

Hand-derived PaiNN forward + backward (forces) written in the SAME dataflow the CUDA kernels
use (surface_sampling_b200/csrc/painn.cu): per-layer stored activations, a message-backward
that is a *gather over the receiver's own CSR row* (each directed edge i<-j is processed
together with its reverse j<-i, so no scatter/atomics are needed), vector features laid out
[A,3,F].  It exists to validate the derivation against autograd (oracle/painn.py) on the CPU
before/alongside the GPU parity tests.  Restates NFF ``Painn`` (see oracle/painn.py header).
"""
from __future__ import annotations

import math

import numpy as np
import torch

from .painn import CUTOFF, FEAT, N_CONV, N_RBF, VEX_POWER, VEX_SIGMA


def _swish(x):
    return x * torch.sigmoid(x)


def _dswish(x):
    sg = torch.sigmoid(x)
    return sg * (1 + x * (1 - sg))


def edge_geometry(xyz, i, j, offsets, cutoff=CUTOFF):
    """Per-edge record shared by all models/layers: unit, d, valid, re=rbf*env, dre=d(re)/dd,
    vex and dvex/dd."""
    r = xyz[j] - xyz[i] + offsets
    d_plain = (r ** 2).sum(-1).sqrt()
    valid = d_plain <= cutoff
    d = ((r ** 2 + 1e-10).sum(-1)).sqrt()
    unit = r / d[:, None]
    n = torch.arange(1, N_RBF + 1, dtype=xyz.dtype)
    coef = n * math.pi / cutoff
    arg = coef[None, :] * d[:, None]
    inside = (d < cutoff)[:, None]
    rbf = torch.where(inside, torch.sin(arg) / d[:, None], torch.zeros_like(arg))
    drbf = torch.where(inside, (coef[None, :] * torch.cos(arg) - torch.sin(arg) / d[:, None]) / d[:, None],
                       torch.zeros_like(arg))
    env = torch.where(d < cutoff, 0.5 * (torch.cos(math.pi * d / cutoff) + 1), torch.zeros_like(d))
    denv = torch.where(d < cutoff, -0.5 * math.pi / cutoff * torch.sin(math.pi * d / cutoff), torch.zeros_like(d))
    re = rbf * env[:, None]
    dre = drbf * env[:, None] + rbf * denv[:, None]
    vex = (VEX_SIGMA / d_plain) ** VEX_POWER
    dvex = -VEX_POWER * vex / d_plain
    return dict(unit=unit, d=d, valid=valid, re=re, dre=dre, env=env, denv=denv, vex=vex, dvex=dvex)


def energy_and_grad_manual(state, pos, numbers, nbr_i, nbr_j, offsets, dtype=torch.float64):
    w = {k: torch.tensor(np.asarray(v), dtype=dtype) for k, v in state.items()}
    xyz = torch.tensor(np.asarray(pos), dtype=dtype)
    z = torch.as_tensor(np.asarray(numbers), dtype=torch.long)
    i = torch.as_tensor(np.asarray(nbr_i), dtype=torch.long)
    j = torch.as_tensor(np.asarray(nbr_j), dtype=torch.long)
    off = torch.tensor(np.asarray(offsets), dtype=dtype)
    A = xyz.shape[0]
    g = edge_geometry(xyz, i, j, off)
    val = g["valid"]
    i, j = i[val], j[val]
    unit, d, re, dre, env, denv = (g[k][val] for k in ("unit", "d", "re", "dre", "env", "denv"))
    vex, dvex = g["vex"][val], g["dvex"][val]

    s = w["embed_block.atom_embed.weight"][z]
    v = torch.zeros(A, 3, FEAT, dtype=dtype)
    saved = []
    for l in range(N_CONV):
        p = f"message_blocks.{l}.inv_message."
        u = f"update_blocks.{l}."
        W1, b1 = w[p + "inv_dense.layers.0.weight"], w[p + "inv_dense.layers.0.bias"]
        W2, b2 = w[p + "inv_dense.layers.1.weight"], w[p + "inv_dense.layers.1.bias"]
        Wd, bd = w[p + "dist_embed.block.1.weight"], w[p + "dist_embed.block.1.bias"]
        h1 = s @ W1.T + b1
        phi = _swish(h1) @ W2.T + b2                     # [A,384]
        wf = re @ Wd.T + env[:, None] * bd[None, :]      # [E,384]  filter (rbf*env folded)
        x = phi[j] * wf
        x0, x1, x2 = x[:, :FEAT], x[:, FEAT:2 * FEAT], x[:, 2 * FEAT:]
        dv_e = x2[:, None, :] * unit[:, :, None] + x0[:, None, :] * v[j]
        s_mid = s + torch.zeros_like(s).index_add_(0, i, x1)
        v_mid = v + torch.zeros_like(v).index_add_(0, i, dv_e)
        U, V = w[u + "u_mat.weight"], w[u + "v_mat.weight"]
        Uv = v_mid @ U.T                                  # [A,3,F]
        Vv = v_mid @ V.T
        nrm = ((Vv ** 2 + 1e-15).sum(1)).sqrt()          # [A,F]
        cat = torch.cat([s_mid, nrm], dim=-1)
        W3, b3 = w[u + "s_dense.0.weight"], w[u + "s_dense.0.bias"]
        W4, b4 = w[u + "s_dense.1.weight"], w[u + "s_dense.1.bias"]
        h3 = cat @ W3.T + b3
        a = _swish(h3) @ W4.T + b4
        a_vv, a_sv, a_ss = a[:, :FEAT], a[:, FEAT:2 * FEAT], a[:, 2 * FEAT:]
        inner = (Uv * Vv).sum(1)
        s_out = s_mid + inner * a_sv + a_ss
        v_out = v_mid + Uv * a_vv[:, None, :]
        saved.append(dict(s_in=s, v_in=v, h1=h1, phi=phi, wf=wf, Uv=Uv, Vv=Vv, nrm=nrm, h3=h3, a=a))
        s, v = s_out, v_out
    r_ = "readout_blocks.0.readoutdict.energy."
    W5, b5, W6, b6 = w[r_ + "0.weight"], w[r_ + "0.bias"], w[r_ + "1.weight"], w[r_ + "1.bias"]
    h5 = s @ W5.T + b5
    e_atom = (_swish(h5) @ W6.T + b6).reshape(-1)
    e_atom = e_atom + torch.zeros_like(e_atom).index_add_(0, i, vex)
    energy = e_atom.sum()

    # ---------------- backward ----------------
    grad = torch.zeros_like(xyz)
    # excluded volume: edge A (in e_i) and its reverse B (in e_j) -> dE/dx_i = -2 dvex * unit
    grad.index_add_(0, i, -2.0 * dvex[:, None] * unit)
    ds = (_dswish(h5) * W6.reshape(1, -1)) @ W5          # [A,F]
    dv = torch.zeros(A, 3, FEAT, dtype=dtype)
    # reverse-edge bookkeeping only for the prototype's vectorised gather: in CUDA thread i
    # simply reads ds[j], dv[j] of the neighbour in its own row.
    for l in reversed(range(N_CONV)):
        sv = saved[l]
        p = f"message_blocks.{l}.inv_message."
        u = f"update_blocks.{l}."
        W1 = w[p + "inv_dense.layers.0.weight"]
        W2 = w[p + "inv_dense.layers.1.weight"]
        Wd, bd = w[p + "dist_embed.block.1.weight"], w[p + "dist_embed.block.1.bias"]
        U, V = w[u + "u_mat.weight"], w[u + "v_mat.weight"]
        W3, W4 = w[u + "s_dense.0.weight"], w[u + "s_dense.1.weight"]
        Uv, Vv, nrm, a, h3 = sv["Uv"], sv["Vv"], sv["nrm"], sv["a"], sv["h3"]
        a_vv, a_sv = a[:, :FEAT], a[:, FEAT:2 * FEAT]
        inner = (Uv * Vv).sum(1)
        # B8: update elementwise backward
        da = torch.cat([(dv * Uv).sum(1), ds * inner, ds], dim=-1)
        dUv = dv * a_vv[:, None, :] + (ds * a_sv)[:, None, :] * Vv
        dVv = (ds * a_sv)[:, None, :] * Uv
        # B7, B6
        dh3 = (da @ W4) * _dswish(h3)
        dcat = dh3 @ W3
        # B5
        ds_mid = ds + dcat[:, :FEAT]
        dVv = dVv + (dcat[:, FEAT:] / nrm)[:, None, :] * Vv
        # B4
        dv_mid = dv + dUv @ U + dVv @ V
        # B3: message backward as a gather over row i (edge A: i<-j, edge B: j<-i, unit_B=-unit)
        phi, v_in, wf = sv["phi"], sv["v_in"], sv["wf"]
        qf = dre @ Wd.T + denv[:, None] * bd[None, :]    # d(filter)/dd  [E,384]
        w0, w1, w2 = wf[:, :FEAT], wf[:, FEAT:2 * FEAT], wf[:, 2 * FEAT:]
        q0, q1, q2 = qf[:, :FEAT], qf[:, FEAT:2 * FEAT], qf[:, 2 * FEAT:]
        gs_i, gv_i, gs_j, gv_j = ds_mid[i], dv_mid[i], ds_mid[j], dv_mid[j]
        phi_i, phi_j = phi[i], phi[j]
        v_i, v_j = v_in[i], v_in[j]
        # edge A
        dxA1 = gs_i
        dxA2 = (gv_i * unit[:, :, None]).sum(1)
        dxA0 = (gv_i * v_j).sum(1)
        dwA0, dwA1, dwA2 = dxA0 * phi_j[:, :FEAT], dxA1 * phi_j[:, FEAT:2 * FEAT], dxA2 * phi_j[:, 2 * FEAT:]
        duA = gv_i * (phi_j[:, 2 * FEAT:] * w2)[:, None, :]          # [E,3,F]
        # edge B
        dxB1 = gs_j
        dxB2 = -(gv_j * unit[:, :, None]).sum(1)
        dxB0 = (gv_j * v_i).sum(1)
        dphi_e = torch.cat([dxB0 * w0, dxB1 * w1, dxB2 * w2], dim=-1)
        dvin_e = (phi_i[:, :FEAT] * w0)[:, None, :] * gv_j
        dwB0, dwB1, dwB2 = dxB0 * phi_i[:, :FEAT], dxB1 * phi_i[:, FEAT:2 * FEAT], dxB2 * phi_i[:, 2 * FEAT:]
        duB = gv_j * (phi_i[:, 2 * FEAT:] * w2)[:, None, :]
        dd = ((dwA0 + dwB0) * q0 + (dwA1 + dwB1) * q1 + (dwA2 + dwB2) * q2)   # [E,F] per-feature
        delta = duA - duB                                                     # [E,3,F]
        proj = (delta * unit[:, :, None]).sum(1)                              # [E,F]
        gpos_e = -(dd[:, None, :] * unit[:, :, None]) - (delta - proj[:, None, :] * unit[:, :, None]) / d[:, None, None]
        grad.index_add_(0, i, gpos_e.sum(-1))
        dphi = torch.zeros_like(phi).index_add_(0, i, dphi_e)
        dv_in = dv_mid + torch.zeros_like(dv_mid).index_add_(0, i, dvin_e)
        # B2, B1
        dh1 = (dphi @ W2) * _dswish(sv["h1"])
        ds = ds_mid + dh1 @ W1
        dv = dv_in
    return energy, grad
