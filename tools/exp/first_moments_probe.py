#!/usr/bin/env python
"""GPU probe: first moments of K x K Legendre bases, register-tiled kernel (default) against the one-function-per-thread kernel
(pyves_set_tuning 96), FP32 and FP64, 10^6 rows."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from pyves_b200 import _lib  # noqa: E402


def t_ms(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
for prec in ("fp32", "fp64"):
    for k in (8, 12, 16, 20, 32, 48, 64):
        e = _lib.Ensemble(n, dims=2, seed=1, precision=prec)
        e.set_integrator(1.0, 2.494, 20.0, 0.01)
        e.set_potential(_lib.POT_DOUBLE_WELL_2D)
        e.set_bias_legendre2d(k - 1, k - 1, (-2.5, 2.5, -2, 2), np.zeros(k * k))
        e.init_positions_mc((-2.5, 2.5, -2, 2), 0.4)
        sw = torch.zeros(1, dtype=torch.float64, device="cuda")
        sf = torch.zeros(k * k, dtype=torch.float64, device="cuda")
        e.set_tuning(95)
        tiled = t_ms(lambda: e.moments(out=(sw, sf, None)))
        a = sf.clone()
        e.set_tuning(96)
        plain = t_ms(lambda: e.moments(out=(sw, sf, None)))
        e.set_tuning(95)
        print(json.dumps(dict(precision=prec, K=k * k, n=n, tiled_ms=tiled, one_function_per_thread_ms=plain,
                              max_abs_diff_over_n=float((a - sf).abs().max()) / n)), flush=True)
        e.close()
