#!/usr/bin/env python
"""GPU run of BASELINE.json configs[1] to convergence: 2D double well, 16x16 Legendre bias, well-tempered
target (gamma 10), averaged SGD; reports the FES RMSE (kT) of the x projection and of the 2D surface.

usage: python tools/converge_cfg2.py [n_walkers] [n_updates] [stride] [mu]
"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyves_b200 import basis, bias, fes, potentials, target  # noqa: E402
from pyves_b200.langevin_dynamics import SingleParticleSimulation  # noqa: E402

potentials._Potential.quiet = True


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
    nup = int(sys.argv[2]) if len(sys.argv) > 2 else 400
    stride = int(sys.argv[3]) if len(sys.argv) > 3 else 500
    mu = float(sys.argv[4]) if len(sys.argv) > 4 else 2.0
    gamma = 10.0
    kT = 8.3145e-3 * 300
    pot = potentials.DoubleWellPotential2D()
    model = basis.LegendreBasis2D(15, 15, -2.5, 2.5, -2.0, 2.0)
    tgt = target.Target_WellTempered_2D(100, 100, gamma)
    decay = float(sys.argv[5]) if len(sys.argv) > 5 else 0.8
    second = len(sys.argv) > 6 and sys.argv[6] == "second"
    b = bias.VESBias_Ensemble(model, tgt, 1 / kT, "averaged_sgd", {"mu": mu, "averaging": "ewma", "decay": decay, "second_order": second},
                              target_update_every=10)
    sim = SingleParticleSimulation(pot, n_walkers=n, seed=20221019)
    sim.place_walkers(xmin=-2.5, xmax=2.5, ymin=-2, ymax=2)
    ens = sim.ens
    b.force.apply(ens)
    xs, F_ref = fes.projection_x(pot, 300, [-2.25, 2.25], [-2, 2])
    hist = []
    t0 = time.time()
    done = 0
    for chunk in range(nup // 50):
        for _ in range(50):
            ens.step(stride)
            sw, sf, sff = ens.moments(cov=second)
            b.update_from_moments(sw, sf, sff)
            b.force.apply(ens)
        done += 50
        # F(x, y) = -V - kT ln p on the target grid, then project on x
        F2 = b.fes().reshape(100, 100) / kT
        Fx = -np.log(np.exp(-(F2 - F2.min())).sum(1))
        cx = -2.5 + 5.0 * (np.arange(100) + 0.5) / 100
        Fx_i = np.interp(xs, cx, Fx - Fx.min())
        rm = fes.rmse_kt(Fx_i, F_ref, clip=20.0)
        gx, gy = np.meshgrid(cx, -2.0 + 4.0 * (np.arange(100) + 0.5) / 100, indexing="ij")
        U = pot.potential(gx, gy) / kT
        U -= U.min()
        m = U <= 20
        d = (F2 - F2[m].min()) - U
        rm2 = float(np.sqrt(np.mean(d[m] ** 2)))
        hist.append(dict(updates=done, md_steps=done * stride, rmse_x_kt=rm, rmse_2d_kt=rm2, wall_s=time.time() - t0))
        print(json.dumps(hist[-1]), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(dict(n_walkers=n, stride=stride, mu=mu, gamma=gamma, history=hist), open("gpurun_out/converge_cfg2.json", "w"), indent=1)


if __name__ == "__main__":
    main()
