"""world_size-2 gloo worker for tests/test_mc_parity.py: rank r owns chains r::world."""
import sys

import numpy as np
import torch
import torch.distributed as dist

from surface_sampling_b200.parallel import gather_chain_stats, shard_chains
from test_mc_parity import _driver

rank, world, out = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
mode = sys.argv[4] if len(sys.argv) > 4 else "chains"
dist.init_process_group("gloo", rank=rank, world_size=world)
if mode == "grid":
    # BASELINE config 5: (pH, U, chain) units of the Pourbaix grid, sharded by parallel.shard_grid (via bench.grid_units)
    from test_mc_parity import _grid_driver
    drv, n_total = _grid_driver(rank, world)
    res = drv.run(total_sweeps=2, sweep_size=3, start_temp=0.257, perform_annealing=False)
    full = gather_chain_stats(res, n_chains_total=n_total, rank=rank, world=world, device="cpu")
else:
    seeds = shard_chains(list(range(6)), rank, world)
    drv, _ = _driver(seeds)
    res = drv.run(total_sweeps=2, sweep_size=4, start_temp=0.5, perform_annealing=False)
    full = gather_chain_stats(res, n_chains_total=6, rank=rank, world=world, device="cpu")
if rank == 0:
    np.save(out + "/gathered.npy", np.stack([full["energy_hist"], full["frac_accept_hist"],
                                             full["adsorption_count_hist"].astype(float)]))
dist.destroy_process_group()
