#!/usr/bin/env python
"""Run a few launches of a light step-kernel configuration (for ncu captures).

usage: python tools/run_light.py [cfg1|cfg3|none] [n=1000000] [steps=500] [launches=3]
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyves_b200 import _lib  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg1"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1000000
S = int(sys.argv[3]) if len(sys.argv) > 3 else 500
L = int(sys.argv[4]) if len(sys.argv) > 4 else 3
rng = np.random.default_rng(0)
e = _lib.Ensemble(n, dims=2, seed=1, precision="fp32")
e.set_integrator(1.0, 8.31446261815324e-3 * 300, 20.0, 0.01)
box = (-2.5, 2.5, -2, 2)
if cfg == "cfg3":
    e.set_potential(_lib.POT_SLIP_BOND_2D, (0, 0, 1, 5, 4, 0, 2))
    e.set_bias_legendre1d(0, 12, -3, 6, rng.standard_normal(13) * 0.1)
    box = (-8, 10, -6, 8)
else:
    e.set_potential(_lib.POT_DOUBLE_WELL_2D, ())
    if cfg == "cfg1":
        e.set_bias_legendre1d(0, 5, -2.5, 2.5, rng.standard_normal(6))
    else:
        e.set_bias_none()
e.init_positions_mc(box, 1 / (8.3145e-3 * 300), ntrials=2000)
e.init_velocities()
for _ in range(L):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    e.step(S)
    b.record()
    torch.cuda.synchronize()
    print("launch: %.3f ms  %.3e walker-steps/s" % (a.elapsed_time(b), n * S / a.elapsed_time(b) * 1e3))
