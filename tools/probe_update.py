#!/usr/bin/env python
"""GPU probe: cost of the update-path kernels (first moments, covariance, MLP parameter-gradient moments)
next to the cost of the MD steps between two updates, per configuration."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyves_b200 import _lib  # noqa: E402


def t_ms(fn, reps=3):
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
    stride = 500
    rng = np.random.default_rng(0)
    rows = []

    def run(tag, pot, params, setup, box, cov):
        e = _lib.Ensemble(n, dims=2, seed=1, precision="fp32")
        e.set_integrator(1.0, 2.494, 20.0, 0.01)
        e.set_potential(pot, params)
        setup(e)
        e.init_positions_mc(box, 0.4)
        e.init_velocities()
        row = dict(tag=tag, n=n, K=e.basis_size, steps_ms=t_ms(lambda: e.step(stride), 2), first_moments_ms=t_ms(lambda: e.moments()))
        if cov:
            row["cov_ms"] = t_ms(lambda: e.moments(cov=True))
            e.set_tuning(98)
            row["cov_cuda_core_ms"] = t_ms(lambda: e.moments(cov=True), 1)
            e.set_tuning(97)
        rows.append(row)
        print(json.dumps(row), flush=True)
        e.close()

    DW, SLIP, SB = _lib.POT_DOUBLE_WELL_2D, _lib.POT_SLIP_BOND_2D, _lib.POT_SZABO_BEREZHKOVSKII
    box = (-2.5, 2.5, -2, 2)
    run("cfg1 leg1d5", DW, (), lambda e: e.set_bias_legendre1d(0, 5, -2.5, 2.5, np.zeros(6)), box, True)
    run("cfg3 slip leg1d12", SLIP, (0, 0, 1, 5, 4, 0, 2), lambda e: e.set_bias_legendre1d(0, 12, -3, 6, np.zeros(13)), (-8, 10, -6, 8), True)
    for k in (8, 16, 32):
        run("leg2d %dx%d" % (k, k), DW, (), lambda e: e.set_bias_legendre2d(k - 1, k - 1, box, np.zeros(k * k)), box, True)
    run("leg2d 64x64", DW, (), lambda e: e.set_bias_legendre2d(63, 63, box, np.zeros(4096)), box, False)
    run("cfg4 mlp 3x48", SB, (2.2, 4.0, 4.04, 4.84), lambda e: e.set_bias_mlp(0, -4, 4, [48, 48, 48], 0, rng.standard_normal(4849) * 0.2), (-4, 4, -4, 4), False)
    run("mlp 2x128", SB, (2.2, 4.0, 4.04, 4.84), lambda e: e.set_bias_mlp(0, -4, 4, [128, 128], 0, rng.standard_normal(128 * 2 + 128 * 128 + 128 + 129) * 0.1), (-4, 4, -4, 4), False)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(rows, open("gpurun_out/probe_update.json", "w"), indent=1)


if __name__ == "__main__":
    main()
