#!/usr/bin/env python
"""All BASELINE.json configurations through the package API on one GPU: walker-steps/s of the full VES
iteration (MD steps + moments + coefficient update + re-upload), one JSON line per configuration.

usage (GPU box): python tools/bench_configs.py [cfg1,cfg2,cfg3,cfg4,cfg5] [iterations]

  cfg1  examples/double_well_2D: LegendreBasis1D(5) along x, STATIC bias (no updates), 2^20 and 2^24 walkers
  cfg2  2D double well, 16 x 16 Legendre, well-tempered target, averaged SGD (bench.py is the reference run)
  cfg3  slip bond, LegendreBasis1D(12) along x, uniform target, 2nd-order averaged SGD (13 x 13 covariance)
  cfg4  Szabo-Berezhkovskii, NNBasis1D [48, 48, 48], Adam 1e-3, uniform target mesh 200, update every 500 steps
  cfg5  sweep: 2D double well, K x K Legendre (K = 8 .. 64), uniform target, 10^5 .. 10^8 walkers
"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyves_b200 import _lib, basis, bias, potentials, target  # noqa: E402

potentials._Potential.quiet = True
KB, KB_PY, TEMP = 8.31446261815324e-3, 8.3145e-3, 300.0
BETA = 1.0 / (KB_PY * TEMP)
STRIDE = 500


def ensemble(n, pot_kind, pot_params, box):
    e = _lib.Ensemble(n, dims=2, seed=20221019, precision="fp32")
    e.set_integrator(1.0, KB * TEMP, 20.0, 0.01)
    e.set_potential(pot_kind, pot_params)
    failed = e.init_positions_mc(box, BETA, ntrials=2000)   # the reference uses 100 for ONE particle (3 % failure on the slip-bond box)
    assert failed == 0, failed
    e.init_velocities()
    return e


def timed(fn, iters, warmup=2):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e-3 / iters


def emit(**row):
    print(json.dumps(row), flush=True)
    return row


def cfg1(iters):
    w = np.array([-11.8, 4.2, 13.1, -2.5, -9.0, 0.4])      # a degree-5 bias of the size the example fits
    for n in (2 ** 20, 2 ** 24):
        e = ensemble(n, _lib.POT_DOUBLE_WELL_2D, (), (-2.5, 2.5, -2, 2))
        b = bias.StaticBias_SingleParticle(basis.LegendreBasis1D(5, -2.5, 2.5, 'x', weights=w), None)
        b.force.apply(e)
        t = timed(lambda: e.step(1000), iters)
        emit(config="cfg1 double well + static LegendreBasis1D(5)", walkers=n, md_steps_per_iteration=1000,
             ms_per_iteration=t * 1e3, walker_steps_per_s=n * 1000 / t)
        e.close()


def cfg2(iters):
    n = 1000000
    e = ensemble(n, _lib.POT_DOUBLE_WELL_2D, (), (-2.5, 2.5, -2, 2))
    b = bias.VESBias_Ensemble(basis.LegendreBasis2D(15, 15, -2.5, 2.5, -2.0, 2.0), target.Target_WellTempered_2D(100, 100, 10.0), BETA,
                              "averaged_sgd", {"mu": 0.5, "averaging": "ewma", "decay": 0.8, "precondition": 1.0}, target_update_every=10)
    upd = bias.DeviceUpdater(b, e)

    def it():
        e.step(STRIDE)
        sw, sf, _ = e.moments()
        upd.update(sw, sf)
    t = timed(it, iters)
    emit(config="cfg2 double well + 16x16 Legendre, well-tempered, averaged SGD (device-resident update)", walkers=n,
         md_steps_per_iteration=STRIDE, ms_per_iteration=t * 1e3, walker_steps_per_s=n * STRIDE / t)
    upd.close()
    e.close()


def cfg3(iters):
    n = 1250000                                              # 10^7 walkers over 8 GPUs
    e = ensemble(n, _lib.POT_SLIP_BOND_2D, (0, 0, 1, 5, 4, 0, 2), (-8, 10, -6, 8))
    b = bias.VESBias_Ensemble(basis.LegendreBasis1D(12, -3, 6, 'x', weights=np.zeros(13)), target.Target_Uniform_HardSwitch_x(200), BETA,
                              "averaged_sgd", {"mu": 0.5, "second_order": True, "averaging": "ewma", "decay": 0.8}, target_update_every=0)
    upd = bias.DeviceUpdater(b, e)

    def it():
        e.step(STRIDE)
        bufs = upd.moment_buffers(True)
        e.moments(cov=True, out=bufs)
        upd.update(*bufs)
    t = timed(it, iters)
    parts = dict(steps_ms=timed(lambda: e.step(STRIDE), 3) * 1e3, moments_cov_ms=timed(lambda: e.moments(cov=True), 3) * 1e3)
    upd.close()
    emit(config="cfg3 slip bond + LegendreBasis1D(12), uniform target, 2nd-order averaged SGD (13x13 covariance, device-resident update)", walkers=n,
         md_steps_per_iteration=STRIDE, ms_per_iteration=t * 1e3, walker_steps_per_s=n * STRIDE / t, **parts)
    e.close()


def cfg4(iters):
    n = 1000000
    e = ensemble(n, _lib.POT_SZABO_BEREZHKOVSKII, (2.2, 4.0, 4.04, 4.84), (-4, 4, -4, 4))
    torch.manual_seed(0)
    model = basis.NNBasis1D(min=-4, max=4, axis='x', hidden_layer_sizes=[48, 48, 48])
    b = bias.VESBias_Ensemble(model, target.Target_Uniform_HardSwitch_x(200), BETA, "Adam", {"lr": 1e-3}, target_update_every=0,
                              gradient_average='running')
    b.force.apply(e)
    upd = bias.DeviceUpdater(b, e)                      # Adam kernel, fold + repack on the device: no host round trip

    def it():
        e.step(STRIDE)
        sw, sf, _ = e.moments()
        upd.update(sw, sf)
    t = timed(it, iters)
    parts = dict(steps_ms=timed(lambda: e.step(STRIDE), 2) * 1e3, moments_ms=timed(lambda: e.moments(), 3) * 1e3)
    upd.close()
    emit(config="cfg4 Szabo-Berezhkovskii + NNBasis1D [48,48,48], Adam 1e-3, running-average gradients (device-resident update)", walkers=n,
         md_steps_per_iteration=STRIDE, ms_per_iteration=t * 1e3, walker_steps_per_s=n * STRIDE / t, **parts)
    e.close()


def cfg5(iters):
    for K in (8, 16, 32, 64):
        for n in (10 ** 5, 10 ** 6, 10 ** 7, 10 ** 8):
            stride = max(10, min(500, int(4e9 / (n * K * K / 64.0))))      # bounded work per point
            e = ensemble(n, _lib.POT_DOUBLE_WELL_2D, (), (-2.5, 2.5, -2, 2))
            b = bias.VESBias_Ensemble(basis.LegendreBasis2D(K - 1, K - 1, -2.5, 2.5, -2.0, 2.0), target.Target_Uniform_2D(100, 100), BETA,
                                      "averaged_sgd", {"mu": 0.5, "averaging": "ewma", "decay": 0.8, "precondition": 1.0}, target_update_every=0)
            upd = bias.DeviceUpdater(b, e)

            def it():
                e.step(stride)
                sw, sf, _ = e.moments()
                upd.update(sw, sf)
            t = timed(it, max(2, iters // 2), warmup=1)
            emit(config="cfg5 double well + %dx%d Legendre, uniform target" % (K, K), walkers=n, md_steps_per_iteration=stride,
                 ms_per_iteration=t * 1e3, walker_steps_per_s=n * stride / t, state_mb=16.0 * n / 1e6)
            upd.close()
            e.close()


if __name__ == "__main__":
    which = sys.argv[1].split(",") if len(sys.argv) > 1 else ["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"]
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    print(json.dumps({"gpu": torch.cuda.get_device_name(0), "fp32_fma_peak_tflops": _lib.fp32_fma_peak(0)}), flush=True)
    for c in which:
        globals()[c](iters)
