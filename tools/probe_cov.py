#!/usr/bin/env python
"""GPU probe: time of the K x K covariance (tcgen05 TF32 vs CUDA cores) for an ensemble of 10^6 walkers."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
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


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
    out = []
    for k in (6, 8, 10, 12, 16, 20, 32):
        e = _lib.Ensemble(n, dims=2, seed=1, precision="fp32")
        e.set_integrator(1.0, 2.494, 20.0, 0.01)
        e.set_potential(_lib.POT_DOUBLE_WELL_2D)
        e.set_bias_legendre2d(k - 1, k - 1, (-2.5, 2.5, -2, 2), np.zeros(k * k))
        e.init_positions_mc((-2.5, 2.5, -2, 2), 0.4)
        K = k * k
        first = t_ms(lambda: e.moments())
        tc = t_ms(lambda: e.moments(cov=True))
        e.set_tuning(98)
        cc = t_ms(lambda: e.moments(cov=True))
        e.set_tuning(97)
        row = dict(K=K, n=n, first_moments_ms=first, cov_default_ms=tc, cov_cuda_core_ms=cc,
                   cov_tflops_default=2.0 * n * K * K / ((tc - first) * 1e-3) * 1e-12 if tc > first else None)
        out.append(row)
        print(json.dumps(row), flush=True)
        e.close()
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/probe_cov.json", "w"), indent=1)


if __name__ == "__main__":
    main()
