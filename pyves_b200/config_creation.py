"""Initial configurations -- mirror of ``ves/config_creation.py`` plus the ensemble version.

``singleParticle2D_init_coord`` keeps the reference semantics (``ves/config_creation.py:41-72``):
up to ``ntrials`` uniform draws from numpy's global generator, accept the first with
``exp(-U / kT) >= 0.5`` (kT = 8.3145e-3 T), z = 0, ``RuntimeError`` otherwise; returns a ``[1, 3]``
array.  ``singleParticle1D_init_coord`` follows ``:10-38`` but calls the 1D potential with one
argument (the reference passes two and raises ``TypeError``).

``ensemble_init_coord`` applies the same acceptance rule to every walker of an ensemble on the GPU
(``pyves_init_positions_mc``), drawing from the Philox placement stream.
"""
import numpy as np

KB_PY = 8.3145 / 1000


def _place(draw, energy, temp, ntrials):
    beta = 1 / (KB_PY * temp)
    for i in range(ntrials):
        trial = draw()
        if np.exp(-beta * energy(trial)) >= 0.5:
            print("[MC particle placement] Particle placement succeeded at iteration {}".format(i))
            return np.array(trial, dtype=float).reshape((1, 3))
    raise RuntimeError("[MC particle placement] Could not place particle on surface.")


def singleParticle1D_init_coord(potential, temp, ntrials=100, xmin=0, xmax=1):
    return _place(lambda: [(xmax - xmin) * np.random.random() + xmin, 0, 0],
                  lambda c: potential.potential(c[0]), temp, ntrials)


def singleParticle2D_init_coord(potential, temp, ntrials=100, xmin=0, xmax=1, ymin=0, ymax=1):
    return _place(lambda: [(xmax - xmin) * np.random.random() + xmin, (ymax - ymin) * np.random.random() + ymin, 0],
                  lambda c: potential.potential(c[0], c[1]), temp, ntrials)


def ensemble_init_coord(ensemble, temp, ntrials=100, xmin=0, xmax=1, ymin=0, ymax=1):
    """Place every walker of ``ensemble`` (a ``pyves_b200._lib.Ensemble`` with its potential set)."""
    failed = ensemble.init_positions_mc((xmin, xmax, ymin, ymax), 1 / (KB_PY * temp), ntrials)
    if failed:
        raise RuntimeError("[MC particle placement] Could not place {} of {} walkers on surface.".format(failed, ensemble.n))
