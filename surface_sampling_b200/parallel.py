"""Multi-GPU: chains (and pH x U grid points) are independent units, so they shard across ranks
with NO data-path collective; torch.distributed (NCCL over NVLink on the B200 box, gloo in the CPU
tests) only gathers the per-chain per-sweep scalars the reference writes to stats.csv
(scripts/sample_surface.py:221-229): surface energy, acceptance fraction, adsorbate count."""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def shard_chains(items: list, rank: int, world: int) -> list:
    """chain c -> rank c mod world."""
    return list(items[rank::world])


def shard_grid(ph_values, u_values, chains_per_point: int, rank: int, world: int):
    """(pH, U, chain) triples round-robin over ranks (BASELINE config 5)."""
    units = [(ph, u, c) for ph in ph_values for u in u_values for c in range(chains_per_point)]
    return units[rank::world]


def gather_chain_stats(res: dict, n_chains_total: int, rank: int, world: int, device="cuda") -> dict:
    """all_gather the [C_local, sweeps] histories and restore global chain order."""
    keys = ("energy_hist", "frac_accept_hist", "adsorption_count_hist")
    if world == 1:
        return {k: np.asarray(res[k]) for k in keys}
    n_sweeps = np.asarray(res["energy_hist"]).shape[1]
    c_max = (n_chains_total + world - 1) // world
    local = np.zeros((3, c_max, n_sweeps))
    for q, k in enumerate(keys):
        a = np.asarray(res[k], dtype=np.float64)
        local[q, :a.shape[0]] = a
    t = torch.from_numpy(local).to(device)
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t)
    full = np.zeros((3, n_chains_total, n_sweeps))
    for r, p in enumerate(parts):
        idx = np.arange(n_chains_total)[r::world]
        full[:, idx] = p.cpu().numpy()[:, :len(idx)]
    return {"energy_hist": full[0], "frac_accept_hist": full[1], "adsorption_count_hist": full[2].astype(int)}
