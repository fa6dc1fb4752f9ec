#!/usr/bin/env python
"""Run a few launches of one step-kernel configuration (for ncu captures).

usage: python tools/run_step.py [basis=16] [variant=-1] [n=1000000] [steps=500] [launches=3]
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyves_b200 import _lib  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 16
variant = int(sys.argv[2]) if len(sys.argv) > 2 else -1
n = int(sys.argv[3]) if len(sys.argv) > 3 else 1000000
S = int(sys.argv[4]) if len(sys.argv) > 4 else 500
L = int(sys.argv[5]) if len(sys.argv) > 5 else 3
rng = np.random.default_rng(0)
e = _lib.Ensemble(n, dims=2, seed=1, precision="fp32")
e.set_integrator(1.0, 8.31446261815324e-3 * 300, 20.0, 0.01)
e.set_potential(_lib.POT_DOUBLE_WELL_2D, ())
e.set_bias_legendre2d(K - 1, K - 1, (-2.5, 2.5, -2, 2), rng.standard_normal(K * K) * 0.1)
e.set_tuning(variant)
e.init_positions_mc((-2.5, 2.5, -2, 2), 1 / (8.3145e-3 * 300))
e.init_velocities()
for _ in range(L):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    e.step(S)
    b.record()
    torch.cuda.synchronize()
    print("launch: %.3f ms  %.3e walker-steps/s" % (a.elapsed_time(b), n * S / a.elapsed_time(b) * 1e3))
