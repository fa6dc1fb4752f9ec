#!/usr/bin/env python
"""GPU run of BASELINE.json configs[1] to convergence: 2D double well, 16x16 Legendre bias, well-tempered
target (gamma 10), averaged SGD; reports the FES RMSE (kT) of the x projection and of the 2D surface.

usage: python tools/converge_cfg2.py [n_walkers] [n_updates] [stride] [mu] [decay] [first|second] [mesh] [timestep_fs] [tag]
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
    mesh = int(sys.argv[7]) if len(sys.argv) > 7 else 100
    tstep = float(sys.argv[8]) if len(sys.argv) > 8 else 10.0
    tag = sys.argv[9] if len(sys.argv) > 9 else "run"
    pre = float(sys.argv[10]) if len(sys.argv) > 10 else 1.0
    tgt = target.Target_WellTempered_2D(mesh, mesh, gamma)
    decay = float(sys.argv[5]) if len(sys.argv) > 5 else 0.8
    second = len(sys.argv) > 6 and sys.argv[6] == "second"
    b = bias.VESBias_Ensemble(model, tgt, 1 / kT, "averaged_sgd", {"mu": mu, "averaging": "ewma", "decay": decay, "second_order": second, "precondition": pre},
                              target_update_every=10)
    scheme = sys.argv[11] if len(sys.argv) > 11 else "openmm_langevin"
    inbox = len(sys.argv) > 12 and sys.argv[12] == "inbox"
    sim = SingleParticleSimulation(pot, n_walkers=n, seed=20221019, timestep=tstep, scheme=scheme)
    sim.place_walkers(xmin=-2.5, xmax=2.5, ymin=-2, ymax=2)
    ens = sim.ens
    b.force.apply(ens)
    xs, F_ref = fes.projection_x(pot, 300, [-2.25, 2.25], [-2, 2])
    hist = []
    t0 = time.time()
    done = 0
    rep = max(50, nup // 10)
    for chunk in range(nup // rep):
        for _ in range(rep):
            ens.step(stride)
            if inbox:
                pos, _ = ens.get_state()
                wrow = ((pos[0].abs() <= 2.5) & (pos[1].abs() <= 2.0)).to(pos.dtype)
                sw, sf, sff = ens.moments(row_weights=wrow, cov=second)
            else:
                sw, sf, sff = ens.moments(cov=second)
            b.update_from_moments(sw, sf, sff)
            b.force.apply(ens)
        done += rep
        # F(x, y) = -V - kT ln p on the target grid, then project on x
        F2 = b.fes().reshape(mesh, mesh) / kT
        Fx = -np.log(np.exp(-(F2 - F2.min())).sum(1))
        cx = -2.5 + 5.0 * (np.arange(mesh) + 0.5) / mesh
        Fx_i = np.interp(xs, cx, Fx - Fx.min())
        rm = fes.rmse_kt(Fx_i, F_ref, clip=20.0)
        gx, gy = np.meshgrid(cx, -2.0 + 4.0 * (np.arange(mesh) + 0.5) / mesh, indexing="ij")
        U = pot.potential(gx, gy) / kT
        U -= U.min()
        m = U <= 20
        d = (F2 - F2[m].min()) - U
        rm2 = float(np.sqrt(np.mean(d[m] ** 2)))
        hist.append(dict(updates=done, md_steps=done * stride, rmse_x_kt=rm, rmse_2d_kt=rm2, wall_s=time.time() - t0))
        print(json.dumps(hist[-1]), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    prof = dict(x=xs[::4].tolist(), F_est=(Fx_i - Fx_i.min())[::4].tolist(), F_ref=F_ref[::4].tolist())
    json.dump(dict(n_walkers=n, stride=stride, mu=mu, gamma=gamma, decay=decay, second_order=second, mesh=mesh, timestep_fs=tstep,
                   history=hist, profile_x=prof), open("gpurun_out/converge_cfg2_%s.json" % tag, "w"), indent=1)


if __name__ == "__main__":
    main()
