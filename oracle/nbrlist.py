"""ORACLE (test infrastructure only — never imported by the product path).

CPU restatement of the directed periodic neighbor list the reference obtains from NFF
``AtomsBatch.update_nbr_list`` / ``torch_nbr_list`` (un-vendored dependency
``nff @ surface-sampling-0.3.0``, pyproject.toml:17).  Reference call sites:
``mcmc/dynamics.py:129`` (``slab.update_nbr_list(update_atoms=True)``) and the flags set in
``mcmc/utils/misc.py:34-42`` (``directed=True, requires_large_offsets=False``).

Semantics restated:
  * directed list: both (i<-j) and (j<-i) are present;
  * PBC by explicit enumeration of lattice translations S (minimum image is wrong for the
    7.87 A SrTiO3 cell at a 6 A radius);
  * membership test in fp32: ``d2 < rc*rc`` and ``d2 != 0`` (upstream uses exactly this mask);
  * column 0 = receiver/centre i, column 1 = sender j, r_ij = x_j - x_i + S.cell.

Because the upstream edge ORDER is unpinned we define it: edges are sorted
lexicographically by (i, j, S0, S1, S2).  The fp32 arithmetic is spelled out so the CUDA
kernel can be bit-exact:

    off_c = fl(fl(fl(S0*a_c) + fl(S1*b_c)) + fl(S2*c_c))
    r_c   = fl(fl(x_jc - x_ic) + off_c)
    d2    = fl(fl(fl(r_x*r_x) + fl(r_y*r_y)) + fl(r_z*r_z))

(no fused multiply-add anywhere).  Positions are used as given (unwrapped); the image range is
derived per pair from fractional coordinates so atoms outside the cell are handled.
"""
from __future__ import annotations

import numpy as np


def cell_heights(cell: np.ndarray) -> np.ndarray:
    """Perpendicular heights h_k = V / |a_{k+1} x a_{k+2}| of a 3x3 row-vector cell (fp64)."""
    cell = np.asarray(cell, dtype=np.float64)
    vol = abs(np.linalg.det(cell))
    h = np.zeros(3)
    for k in range(3):
        cr = np.cross(cell[(k + 1) % 3], cell[(k + 2) % 3])
        h[k] = vol / np.linalg.norm(cr)
    return h


def image_bounds(pos: np.ndarray, cell: np.ndarray, pbc, rc: float) -> np.ndarray:
    """Per-structure bound n_k on |S_k| that is a superset of all valid images."""
    cell = np.asarray(cell, dtype=np.float64)
    frac = np.asarray(pos, dtype=np.float64) @ np.linalg.inv(cell)
    h = cell_heights(cell)
    n = np.zeros(3, dtype=np.int64)
    for k in range(3):
        if pbc[k]:
            spread = frac[:, k].max() - frac[:, k].min() if len(frac) else 0.0
            n[k] = int(np.floor(rc / h[k] + spread)) + 1
    return n


def neighbor_list(pos, cell, pbc, rc):
    """Directed neighbor list.

    Args:
        pos: [N,3] positions (cast to fp32), cell: [3,3] row vectors (cast to fp32),
        pbc: 3 bools, rc: cutoff radius (fp32).
    Returns:
        i [E] int32 (receiver), j [E] int32 (sender), S [E,3] int32 lattice shifts, sorted by
        (i, j, S0, S1, S2).
    """
    x = np.ascontiguousarray(pos, dtype=np.float32)
    c = np.ascontiguousarray(cell, dtype=np.float32)
    n_atoms = x.shape[0]
    rc32 = np.float32(rc)
    rc2 = np.float32(rc32 * rc32)
    nb = image_bounds(x.astype(np.float64), c.astype(np.float64), pbc, float(rc))
    s0 = np.arange(-nb[0], nb[0] + 1)
    s1 = np.arange(-nb[1], nb[1] + 1)
    s2 = np.arange(-nb[2], nb[2] + 1)
    S = np.stack(np.meshgrid(s0, s1, s2, indexing="ij"), axis=-1).reshape(-1, 3)  # lexicographic
    Sf = S.astype(np.float32)
    # off_c = ((S0*a_c) + (S1*b_c)) + (S2*c_c), all fp32, no FMA
    off = (Sf[:, 0:1] * c[0][None, :] + Sf[:, 1:2] * c[1][None, :]) + Sf[:, 2:3] * c[2][None, :]
    off = off.astype(np.float32)
    ii, jj, ss = [], [], []
    for i in range(n_atoms):
        dx = (x - x[i][None, :]).astype(np.float32)  # [N,3]  x_j - x_i
        r = (dx[:, None, :] + off[None, :, :]).astype(np.float32)  # [N,nS,3]
        sq = (r * r).astype(np.float32)
        d2 = ((sq[..., 0] + sq[..., 1]).astype(np.float32) + sq[..., 2]).astype(np.float32)
        mask = (d2 < rc2) & (d2 != 0)
        jidx, sidx = np.nonzero(mask)  # row-major: sorted by j then S index (lexicographic)
        ii.append(np.full(jidx.shape, i, dtype=np.int32))
        jj.append(jidx.astype(np.int32))
        ss.append(S[sidx].astype(np.int32))
    if n_atoms == 0:
        return (np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros((0, 3), np.int32))
    return np.concatenate(ii), np.concatenate(jj), np.concatenate(ss, axis=0)


def to_csr(i: np.ndarray, n_atoms: int) -> np.ndarray:
    """Row pointer [N+1] for a receiver-sorted edge list."""
    cnt = np.bincount(i, minlength=n_atoms)
    rp = np.zeros(n_atoms + 1, dtype=np.int32)
    rp[1:] = np.cumsum(cnt)
    return rp
