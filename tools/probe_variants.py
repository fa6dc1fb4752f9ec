#!/usr/bin/env python
"""GPU probe: walker-steps/s of the step kernel for each configuration / kernel variant.

usage (on the GPU box): python tools/probe_variants.py [n_walkers] [steps_per_launch]
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyves_b200 import _lib  # noqa: E402

KB = 8.31446261815324e-3


def time_steps(e, nsteps, reps=3):
    e.step(nsteps)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        e.step(nsteps)
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) * 1e-3)
    return best


def make(n, prec, pot, params=()):
    e = _lib.Ensemble(n, dims=2, seed=1, precision=prec)
    e.set_integrator(1.0, KB * 300, 20.0, 0.01)
    e.set_potential(pot, params)
    return e


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
    S = int(sys.argv[2]) if len(sys.argv) > 2 else 500
    rng = np.random.default_rng(0)
    out = {"n": n, "steps_per_launch": S, "fp32_fma_peak_tflops": _lib.fp32_fma_peak(0)}
    out["fp32_fma2_peak_tflops"] = _lib.fp32_fma_peak(0, packed=True)
    print("FP32 FMA peak: %.1f TFLOP/s (FFMA), %.1f TFLOP/s (FFMA2)" % (out["fp32_fma_peak_tflops"], out["fp32_fma2_peak_tflops"]), flush=True)
    rows = []

    def run(tag, prec, pot, params, setup, variants=(-1,), flops=None, box=(-2.5, 2.5, -2, 2)):
        for v in variants:
            e = make(n, prec, pot, params)
            setup(e)
            e.set_tuning(v)
            e.init_positions_mc(box, 1 / (8.3145e-3 * 300))
            e.init_velocities()
            t = time_steps(e, S)
            rate = n * S / t
            row = dict(tag=tag, prec=prec, variant=v, ms=t * 1e3, walker_steps_per_s=rate)
            if flops:
                row["tflops"] = rate * flops * 1e-12
                row["frac_of_measured_fma_peak"] = row["tflops"] / out["fp32_fma_peak_tflops"]
            rows.append(row)
            print(json.dumps(row), flush=True)
            e.close()

    DW, SLIP, SB = _lib.POT_DOUBLE_WELL_2D, _lib.POT_SLIP_BOND_2D, _lib.POT_SZABO_BEREZHKOVSKII
    slip = (0, 0, 1, 5, 4, 0, 2)
    sb = (2.2, 4.0, 4.04, 4.84)
    only = sys.argv[3].split(",") if len(sys.argv) > 3 else None
    _run = run

    def run(tag, *a, **k):
        if only is None or any(o in tag for o in only):
            _run(tag, *a, **k)

    for prec in ("fp32", "fp64") if (len(sys.argv) <= 4) else (sys.argv[4],):
        run("dw_none", prec, DW, (), lambda e: e.set_bias_none(), variants=(-1, 1), flops=32)
        run("cfg1_dw_leg1d5", prec, DW, (), lambda e: e.set_bias_legendre1d(0, 5, -2.5, 2.5, rng.standard_normal(6)), variants=(-1, 1), flops=70)
        run("cfg3_slip_leg1d12", prec, SLIP, slip, lambda e: e.set_bias_legendre1d(0, 12, -3, 6, rng.standard_normal(13) * 0.1), variants=(-1, 1), flops=133, box=(-8, 10, -6, 8))
        run("dw_leg2d8", prec, DW, (), lambda e: e.set_bias_legendre2d(7, 7, (-2.5, 2.5, -2, 2), rng.standard_normal(64) * 0.1), variants=(-1, 1, 2), flops=4 * 64 + 32 + 84 + 32)
        run("cfg2_dw_leg2d16", prec, DW, (), lambda e: e.set_bias_legendre2d(15, 15, (-2.5, 2.5, -2, 2), rng.standard_normal(256) * 0.1),
            variants=(-1, 1, 2), flops=1300)
        run("cfg5_dw_leg2d32", prec, DW, (), lambda e: e.set_bias_legendre2d(31, 31, (-2.5, 2.5, -2, 2), rng.standard_normal(1024) * 0.05), flops=4628)
        run("cfg5_dw_leg2d64", prec, DW, (), lambda e: e.set_bias_legendre2d(63, 63, (-2.5, 2.5, -2, 2), rng.standard_normal(4096) * 0.02), flops=17428)
        run("cfg4_sb_mlp48x3", prec, SB, sb, lambda e: e.set_bias_mlp(0, -4, 4, [48, 48, 48], 0, rng.standard_normal(4849) * 0.2), flops=9635, box=(-4, 4, -4, 4))
        run("sb_mlp128x2", prec, SB, sb, lambda e: e.set_bias_mlp(0, -4, 4, [128, 128], 0, rng.standard_normal(128 * 2 + 128 * 128 + 128 + 129) * 0.1), box=(-4, 4, -4, 4))
    out["rows"] = rows
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/probe_variants.json", "w") as fh:
        json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()
