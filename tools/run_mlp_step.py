#!/usr/bin/env python
"""Run a few launches of the NN-bias step kernel (for ncu captures): python tools/run_mlp_step.py [variant=-1] [n=1000000] [steps=500] [launches=3]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyves_b200 import _lib, basis  # noqa: E402

variant = int(sys.argv[1]) if len(sys.argv) > 1 else -1
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1000000
S = int(sys.argv[3]) if len(sys.argv) > 3 else 500
L = int(sys.argv[4]) if len(sys.argv) > 4 else 3
torch.manual_seed(0)
m = basis.NNBasis1D(min=-4, max=4, axis="x", hidden_layer_sizes=[48, 48, 48])
theta = np.concatenate([p.detach().numpy().ravel() for p in m.parameters()])
e = _lib.Ensemble(n, dims=2, seed=1, precision="fp32")
e.set_integrator(1.0, 8.31446261815324e-3 * 300, 20.0, 0.01)
e.set_potential(_lib.POT_SZABO_BEREZHKOVSKII, (2.2, 4.0, 4.04, 4.84))
e.set_bias_mlp(0, -4.0, 4.0, [48, 48, 48], 0, theta)
e.set_tuning(variant)
e.init_positions_mc((-4, 4, -4, 4), 1 / (8.3145e-3 * 300), ntrials=2000)
e.init_velocities()
for _ in range(L):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    e.step(S)
    b.record()
    torch.cuda.synchronize()
    print("launch: %.3f ms  %.3e walker-steps/s" % (a.elapsed_time(b), n * S / a.elapsed_time(b) * 1e3))
