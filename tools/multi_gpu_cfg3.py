#!/usr/bin/env python
"""configs[2] shape on N GPUs through the driver (torchrun): slip bond, LegendreBasis1D(12), uniform target,
second-order averaged SGD, walkers sharded by global id, one NCCL all-reduce of [sum_w, sum_f, sum_ff] per update,
device-resident optimiser.  Prints walker-steps/s of the whole run and a checksum of the coefficients per rank
(identical on every rank: every rank applies the same deterministic update).

    torchrun --nproc-per-node N tools/multi_gpu_cfg3.py [walkers_total=2500000] [nsteps=20000]
"""
import contextlib
import io
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyves_b200 import basis, bias, potentials, target  # noqa: E402
from pyves_b200.langevin_dynamics import SingleParticleSimulation  # noqa: E402

potentials._Potential.quiet = True
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n_total = int(sys.argv[1]) if len(sys.argv) > 1 else 2500000
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
kT = 8.3145e-3 * 300
model = basis.LegendreBasis1D(12, -12.0, 14.0, 'x', weights=np.zeros(13))
b = bias.VESBias_Ensemble(model, target.Target_Uniform_HardSwitch_x(801), 1 / kT, "averaged_sgd",
                          {"mu": 3.0, "second_order": True, "averaging": "ewma", "decay": 0.9})
with contextlib.redirect_stdout(io.StringIO()):
    sim = SingleParticleSimulation(potentials.SlipBondPotential2D(), n_walkers=n_total, seed=3, device=local, device_update=True)
    sim.place_walkers(ntrials=2000, xmin=-8, xmax=10, ymin=-6, ymax=8)
    sim.init_ves(b, startafter=500, learnevery=500)
    if world > 1:                                   # create the NCCL communicator outside the timed region
        dist.all_reduce(torch.zeros(1, device="cuda"))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sim(nsteps=nsteps + 1, chkevery=0, trajevery=0, energyevery=0, chkfile=None, trajfile=None, energyfile=None)
    torch.cuda.synchronize()
dt = time.perf_counter() - t0
chk = float(np.sum(b.optimizer.abar * np.arange(1, 14)))
print("rank %d/%d: %d walkers local, %d updates, %.3e walker-steps/s (whole job), coefficient checksum %.12f"
      % (rank, world, sim.n_walkers, b.n_updates, n_total * nsteps / dt, chk), flush=True)
if world > 1:
    dist.destroy_process_group()
