#!/usr/bin/env python
"""GPU probe: time of the K x K covariance (tcgen05 path vs CUDA cores) for an ensemble of n walkers.

Reported rates, both for the covariance launches alone (first-moment kernels timed separately and subtracted):
  executed    = tiles x 2 x 256^2 x n flop (what the tensor core is asked to do: upper-triangular 256 x 256 tiles, the
                lower-left quarter of a diagonal tile skipped)
  algorithmic = n K (K + 1) flop (the upper triangle of C)
against the dense FP16 peak of MEASURED_PEAKS.json (the pipe the kernel uses) and its half (the TF32 rate the
north-star names).
"""
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
    ks = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [8, 10, 12, 16, 20, 32, 48, 64]
    peak16 = 1643.1
    try:
        peak16 = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["bf16_tflops"])
    except (OSError, KeyError, ValueError):
        pass
    out = []
    for k in ks:
        e = _lib.Ensemble(n, dims=2, seed=1, precision="fp32")
        e.set_integrator(1.0, 2.494, 20.0, 0.01)
        e.set_potential(_lib.POT_DOUBLE_WELL_2D)
        e.set_bias_legendre2d(k - 1, k - 1, (-2.5, 2.5, -2, 2), np.zeros(k * k))
        e.init_positions_mc((-2.5, 2.5, -2, 2), 0.4)
        K = k * k
        sw = torch.zeros(1, dtype=torch.float64, device="cuda")
        sf = torch.zeros(K, dtype=torch.float64, device="cuda")
        sff = torch.zeros((K, K), dtype=torch.float64, device="cuda")
        first = t_ms(lambda: e.moments(out=(sw, sf, None)))
        tc = t_ms(lambda: e.moments(out=(sw, sf, sff)))
        w = torch.rand(n, dtype=torch.float32, device="cuda")
        pos = e.get_state()[0].T.contiguous()
        firstw = t_ms(lambda: e.moments(pos=pos, row_weights=w, out=(sw, sf, None)))
        tcw = t_ms(lambda: e.moments(pos=pos, row_weights=w, out=(sw, sf, sff)))
        cc = None
        if K <= 1024:
            e.set_tuning(98)
            cc = t_ms(lambda: e.moments(out=(sw, sf, sff)), reps=2)
            e.set_tuning(97)
        T = (K + 255) // 256
        tiles = T * (T + 1) // 2
        executed = (tiles * 2.0 - T * 0.5) * 256 * 256 * n          # a diagonal tile runs 3/4 of its MMAs
        algorithmic = float(n) * K * (K + 1)
        dt = (tc - first) * 1e-3
        row = dict(K=K, n=n, first_moments_ms=first, cov_total_ms=tc, cov_only_ms=tc - first, cov_weighted_only_ms=tcw - firstw, cov_cuda_core_ms=cc,
                   executed_tflops=executed / dt * 1e-12, algorithmic_tflops=algorithmic / dt * 1e-12,
                   frac_fp16_peak_executed=executed / dt * 1e-12 / peak16, frac_tf32_peak_executed=executed / dt * 1e-12 / (peak16 / 2),
                   frac_tf32_peak_algorithmic=algorithmic / dt * 1e-12 / (peak16 / 2), fp16_peak_tflops=peak16)
        out.append(row)
        print(json.dumps(row), flush=True)
        e.close()
        del sff
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/probe_cov.json", "w"), indent=1)


if __name__ == "__main__":
    main()
