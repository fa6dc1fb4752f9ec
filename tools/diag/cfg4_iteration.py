#!/usr/bin/env python
"""A few cfg4 iterations (ReLU NNBasis1D [48,48,48], Adam, running-average gradients, device-resident update) for an ncu launch list."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from pyves_b200 import _lib, basis, bias, target
n, stride, iters = 1000000, 500, int(sys.argv[1]) if len(sys.argv) > 1 else 6
torch.manual_seed(0)
b = bias.VESBias_Ensemble(basis.NNBasis1D(min=-4, max=4, axis='x', hidden_layer_sizes=[48, 48, 48]), target.Target_Uniform_HardSwitch_x(200),
                          1.0 / (8.3145e-3 * 300), "Adam", {"lr": 1e-3}, target_update_every=0, gradient_average='running')
e = _lib.Ensemble(n, dims=2, seed=1, precision="fp32")
e.set_integrator(1.0, 8.31446261815324e-3 * 300, 20.0, 0.01)
e.set_potential(_lib.POT_SZABO_BEREZHKOVSKII, (2.2, 4.0, 4.04, 4.84))
e.init_positions_mc((-4, 4, -4, 4), 1.0 / (8.3145e-3 * 300), ntrials=2000)
e.init_velocities()
upd = bias.DeviceUpdater(b, e)
for _ in range(iters):
    e.step(stride)
    bufs = upd.moment_buffers(False)
    e.moments(out=bufs)
    upd.update(*bufs)
torch.cuda.synchronize()
print("done")
