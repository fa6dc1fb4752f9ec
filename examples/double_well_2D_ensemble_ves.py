#!/usr/bin/env python
"""BASELINE.json configs[1] through the reference-style API: an ensemble of walkers on the 2D double well,
16 x 16 Legendre bias, well-tempered target, averaged SGD with the update loop resident on the GPU.

    python examples/double_well_2D_ensemble_ves.py [n_walkers=200000] [n_steps=60000]

The script body follows examples/double_well_2D/static_legendre_double_well.py:113-133 of apallath/pyves
(potential -> simulation -> init_ves -> call); what is new is n_walkers, the 2D basis, the ensemble bias and
the target.  Prints the RMSE (kT) of the converged free energy surface against the analytic one.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pyves_b200  # noqa: E402

pyves_b200.install_as_ves()                     # the package answers to `import ves...` from here on

from ves.basis import LegendreBasis2D  # noqa: E402
from ves.bias import VESBias_Ensemble  # noqa: E402
from ves.langevin_dynamics import SingleParticleSimulation  # noqa: E402
from ves.potentials import DoubleWellPotential2D  # noqa: E402
from ves.target import Target_WellTempered_2D  # noqa: E402


def main(n_walkers=200000, n_steps=60000, quiet=False):
    temp = 300
    beta = 1 / (8.3145e-3 * temp)
    pot = DoubleWellPotential2D()
    V_module = LegendreBasis2D(15, 15, -2.5, 2.5, -2.0, 2.0)
    target = Target_WellTempered_2D(100, 100, bias_factor=10.0)
    ves_bias = VESBias_Ensemble(V_module, target, beta, "averaged_sgd",
                                {"mu": 0.5, "averaging": "ewma", "decay": 0.8, "precondition": 1.0},
                                model_loc=None, target_update_every=10)
    sim = SingleParticleSimulation(pot, temp=temp, n_walkers=n_walkers, seed=20221019, device_update=True)
    sim.place_walkers(xmin=-2.5, xmax=2.5, ymin=-2, ymax=2)          # ves.config_creation rule, every walker
    sim.init_ves(ves_bias, startafter=100, learnevery=100)
    sim(nsteps=n_steps + 1, chkevery=0, trajevery=0, energyevery=0, chkfile=None, trajfile=None, energyfile=None)

    kT = 1 / beta
    F = ves_bias.fes().reshape(target.mesh) / kT                      # F = -V - kT ln p on the target grid
    cx = -2.5 + 5.0 * (np.arange(100) + 0.5) / 100
    cy = -2.0 + 4.0 * (np.arange(100) + 0.5) / 100
    gx, gy = np.meshgrid(cx, cy, indexing="ij")
    U = pot.potential(gx, gy) / kT
    U -= U.min()
    sel = U <= 20
    rmse = float(np.sqrt(np.mean(((F - F[sel].min()) - U)[sel] ** 2)))
    if not quiet:
        print("updates: %d   FES RMSE over the region F <= 20 kT: %.3f kT" % (ves_bias.n_updates, rmse))
    return rmse


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 200000, int(sys.argv[2]) if len(sys.argv) > 2 else 60000)
