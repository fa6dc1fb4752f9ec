#!/usr/bin/env python
"""Experiment: per-step contraction of the 2D Legendre bias on the tensor cores (pyves_set_tuning(h, 7), csrc/langevin_tc.cu)
against the CUDA-core step kernels: parity with an identical injected noise stream, then walker-steps/s."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from pyves_b200 import _lib  # noqa: E402

KB = 8.31446261815324e-3
MM = (-2.5, 2.5, -2.0, 2.0)


def ens(n, K, w, variant, seed=1):
    e = _lib.Ensemble(n, dims=2, seed=seed, precision="fp32")
    e.set_integrator(1.0, KB * 300, 20.0, 0.01)
    e.set_potential(_lib.POT_DOUBLE_WELL_2D)
    e.set_bias_legendre2d(K - 1, K - 1, MM, w)
    e.set_tuning(variant)
    return e


def main():
    sizes = [int(v) for v in sys.argv[1].split(",")] if len(sys.argv) > 1 else [64, 32, 16]
    n_time = int(sys.argv[2]) if len(sys.argv) > 2 else 1000000
    rng = np.random.default_rng(0)
    for K in sizes:
        w = rng.standard_normal((K, K)) * 3.0 / (1.0 + np.add.outer(np.arange(K), np.arange(K)))
        # ---- parity: injected noise, 60 steps, walkers spread over the whole box (and a few outside)
        n, nsteps = 5000, 60
        x0 = np.stack([rng.uniform(-2.7, 2.7, n), rng.uniform(-2.2, 2.2, n)]).astype(np.float32)
        v0 = (rng.standard_normal((2, n)) * 1.58).astype(np.float32)
        noise = torch.tensor(rng.standard_normal((nsteps, 2, n)), dtype=torch.float32, device="cuda")
        res = {}
        for variant in (-1, 7):
            e = ens(n, K, w, variant)
            e.set_state(x0, v0)
            e.step(nsteps, noise)
            res[variant] = [t.copy() for t in e.get_state(host=True)]
            e.close()
        dx = float(np.max(np.abs(res[7][0] - res[-1][0])))
        dv = float(np.max(np.abs(res[7][1] - res[-1][1])))
        # one step: force difference relative to the largest force
        f = {}
        for variant in (-1, 7):
            e = ens(n, K, w, variant)
            e.set_state(x0, np.zeros_like(v0))
            e.step(1, torch.zeros((1, 2, n), dtype=torch.float32, device="cuda"))
            f[variant] = e.get_state(host=True)[1].copy()      # v after one noiseless step from rest = -b F-gradient
            e.close()
        df = float(np.max(np.abs(f[7] - f[-1])) / np.max(np.abs(f[-1])))
        # ---- Philox path: identical stream?
        ph = {}
        for variant in (-1, 7):
            e = ens(n, K, w, variant, seed=7)
            e.set_state(x0, v0)
            e.step(31)
            e.step(30)
            ph[variant] = e.get_state(host=True)[0].copy()
            e.close()
        dph = float(np.max(np.abs(ph[7] - ph[-1])))
        # ---- timing
        out = dict(K=K, parity_max_dx=dx, parity_max_dv=dv, one_step_force_rel=df, philox_max_dx=dph)
        for variant in (-1, 7):
            e = ens(n_time, K, w, variant)
            assert e.init_positions_mc(MM, 0.4) == 0
            e.init_velocities()
            steps = max(20, int(2.0e10 / (n_time * K * K / 16.0)))
            e.step(steps)
            torch.cuda.synchronize()
            best = 1e30
            for _ in range(3):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                e.step(steps)
                b.record()
                torch.cuda.synchronize()
                best = min(best, a.elapsed_time(b))
            out["steps"] = steps
            out["ms_cuda_core" if variant == -1 else "ms_tensor_core"] = best
            out["rate_cuda_core" if variant == -1 else "rate_tensor_core"] = n_time * steps / best * 1e3
            e.close()
        out["speedup"] = out["ms_cuda_core"] / out["ms_tensor_core"]
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
