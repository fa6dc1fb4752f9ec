"""Oracle (test infrastructure): the Langevin step of ``openmm.LangevinIntegrator``.

PARITY UNPINNED: OpenMM is an un-vendored dependency (``requirements.txt:3``,
``OpenMM>=7.6.0``, no lock) that is absent from this image, so this is a restatement of its
published ``ReferenceStochasticDynamics`` scheme, anchored on the
reference's call sites: construction ``ves/langevin_dynamics.py:49-51`` (temperature,
friction, timestep), defaults ``:23-26`` (m = 1 Da, T = 300 K, gamma = 20 / ps, dt = 10 fs),
``setVelocitiesToTemperature`` ``:70-74`` and ``integrator.step(1)`` ``:188``.

    a = exp(-gamma dt);  b = (1 - a) / gamma  (dt if gamma == 0);  c = sqrt(kT (1 - a^2))
    v <- a v + b F / m + c xi / sqrt(m);     x <- x + dt v

The optional ``middle`` scheme is OpenMM's ``LangevinMiddleIntegrator`` (not used by the
reference; offered as ``scheme="middle"``):
    v += dt F/m;  x += dt/2 v;  v <- a v + sqrt(kT (1 - a^2) / m) xi;  x += dt/2 v
"""
import math

import numpy as np

from . import philox

KB = 8.31446261815324e-3        # kJ/mol/K, the constant OpenMM uses
KB_PY = 8.3145e-3               # the constant the reference's Python side uses

SCHEME_OPENMM = 0
SCHEME_MIDDLE = 1


def constants(mass=1.0, temp=300.0, friction=20.0, timestep_fs=10.0):
    """Per-integrator constants in nm / ps / (kJ/mol) units."""
    dt = timestep_fs * 1e-3
    kT = KB * temp
    a = math.exp(-friction * dt)
    b = (1.0 - a) / friction if friction != 0.0 else dt
    c = math.sqrt(kT * (1.0 - a * a))
    return dict(mass=mass, kT=kT, dt=dt, friction=friction, a=a, b=b, c=c,
                b_over_m=b / mass, c_over_sqrtm=c / math.sqrt(mass),
                sigma_v=math.sqrt(kT / mass))


def step(x, v, force_fn, k, xi, scheme=SCHEME_OPENMM):
    """One step for positions/velocities [N, D]; ``force_fn(x) -> F [N, D]``; ``xi`` [N, D]."""
    if scheme == SCHEME_OPENMM:
        F = force_fn(x)
        v = k['a'] * v + k['b_over_m'] * F + k['c_over_sqrtm'] * xi
        x = x + k['dt'] * v
        return x, v
    if scheme == SCHEME_MIDDLE:
        F = force_fn(x)
        v = v + k['dt'] / k['mass'] * F
        x = x + 0.5 * k['dt'] * v
        v = k['a'] * v + k['c_over_sqrtm'] * xi
        x = x + 0.5 * k['dt'] * v
        return x, v
    raise ValueError("unknown scheme")


def run_injected(x, v, force_fn, k, noise, scheme=SCHEME_OPENMM, record=False):
    """``noise`` [nsteps, D, N] (the layout the CUDA kernel reads).  Returns x, v (and frames
    before each step, as ``ves/langevin_dynamics.py:181`` writes them, when ``record``)."""
    x = np.array(x, dtype=np.float64)
    v = np.array(v, dtype=np.float64)
    frames = []
    for t in range(noise.shape[0]):
        if record:
            frames.append(x.copy())
        x, v = step(x, v, force_fn, k, noise[t].T, scheme)
    return (x, v, np.array(frames)) if record else (x, v)


def run_philox(x, v, force_fn, k, seed, gid0, step0, nsteps, scheme=SCHEME_OPENMM):
    """Steps driven by the framework's Philox stream (global walker ids gid0 .. gid0+N-1)."""
    x = np.array(x, dtype=np.float64)
    v = np.array(v, dtype=np.float64)
    gids = np.arange(gid0, gid0 + x.shape[0], dtype=np.uint64)
    for t in range(step0, step0 + nsteps):
        xi = philox.dynamics_noise(seed, gids, t, dims=x.shape[1])
        x, v = step(x, v, force_fn, k, xi, scheme)
    return x, v


def maxwell_boltzmann(seed, gid0, n, k, dims):
    """v = sqrt(kT/m) N(0,1) per component (``ves/langevin_dynamics.py:70-74``)."""
    gids = np.arange(gid0, gid0 + n, dtype=np.uint64)
    return k['sigma_v'] * philox.velocity_noise(seed, gids)[:, :dims]


def kinetic_energy(v, F, k):
    """OpenMM reports the leap-frog kinetic energy at the half-step-shifted velocity
    v + dt/2 F/m [UPSTREAM]; per walker, summed over dimensions."""
    vs = v + 0.5 * k['dt'] * F / k['mass']
    return 0.5 * k['mass'] * np.sum(vs * vs, axis=1)


def mc_placement(seed, gid0, n, energy_fn, beta, box, ntrials=100):
    """Monte-Carlo placement rule of ``ves/config_creation.py:56-68``: up to ``ntrials``
    uniform draws in ``box = (xmin, xmax, ymin, ymax)``; accept the first with
    exp(-beta U) >= 0.5; z = 0.  Returns positions [n, 2] and a success mask."""
    gids = np.arange(gid0, gid0 + n, dtype=np.uint64)
    pos = np.zeros((n, 2))
    done = np.zeros(n, dtype=bool)
    for trial in range(ntrials):
        ux, uy = philox.placement_uniforms(seed, gids, trial)
        tx = box[0] + (box[1] - box[0]) * ux
        ty = box[2] + (box[3] - box[2]) * uy
        ok = np.exp(-beta * energy_fn(tx, ty)) >= 0.5
        take = ok & ~done
        pos[take, 0] = tx[take]
        pos[take, 1] = ty[take]
        done |= take
        if done.all():
            break
    return pos, done
