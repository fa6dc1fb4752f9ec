#!/usr/bin/env python
"""Run a few covariance launches of one basis size (for ncu captures): python tools/run_cov.py [k=32] [n=1000000] [launches=3] [weighted=0]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyves_b200 import _lib  # noqa: E402

k = int(sys.argv[1]) if len(sys.argv) > 1 else 32
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1000000
L = int(sys.argv[3]) if len(sys.argv) > 3 else 3
weighted = int(sys.argv[4]) if len(sys.argv) > 4 else 0
e = _lib.Ensemble(n, dims=2, seed=1, precision="fp32")
e.set_integrator(1.0, 2.494, 20.0, 0.01)
e.set_potential(_lib.POT_DOUBLE_WELL_2D)
e.set_bias_legendre2d(k - 1, k - 1, (-2.5, 2.5, -2, 2), np.zeros(k * k))
e.init_positions_mc((-2.5, 2.5, -2, 2), 0.4)
K = k * k
out = (torch.zeros(1, dtype=torch.float64, device="cuda"), torch.zeros(K, dtype=torch.float64, device="cuda"), torch.zeros((K, K), dtype=torch.float64, device="cuda"))
w = torch.rand(n, dtype=torch.float32, device="cuda") if weighted else None
pos = e.get_state()[0].T.contiguous() if weighted else None
for _ in range(L):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    e.moments(pos=pos, row_weights=w, out=out)
    b.record()
    torch.cuda.synchronize()
    print("moments + covariance K = %d, n = %d: %.3f ms" % (K, n, a.elapsed_time(b)))
