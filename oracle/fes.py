"""Oracle (test infrastructure): free-energy surfaces.

* ``projection_x`` / ``projection_y`` follow ``ves/visualization.py:185-201`` and ``:221-237``:
  U on a mesh x mesh grid (``np.meshgrid`` of two linspaces), divided by kT = 8.3145e-3 T
  (``:137``), ``F = -logsumexp(-V, axis)``, shifted to min 0.  Values in kT.
* ``histogram_fes_2d`` follows ``ves/visualization.py:475-478``: 2D histogram, empty bins take
  the smallest non-zero count, ``beta F = -log counts``, shifted to min 0.
* ``histogram_fes_1d`` is the same rule in one dimension (used for the FES-RMSE metric).
"""
import numpy as np

from . import potentials

KB_PY = 8.3145e-3


def _lse(a, axis):
    m = a.max(axis=axis, keepdims=True)
    return (m + np.log(np.exp(a - m).sum(axis=axis, keepdims=True))).squeeze(axis)


def grid_energy(kind, xrange, yrange, mesh=200, params=None, bias_fn=None):
    xx, yy = np.meshgrid(np.linspace(xrange[0], xrange[1], mesh),
                         np.linspace(yrange[0], yrange[1], mesh))
    v, _, _ = potentials.energy_grad_xy(kind, xx.ravel(), yy.ravel(), params)
    if bias_fn is not None:
        v = v + bias_fn(np.stack([xx.ravel(), yy.ravel()], axis=1))
        v = v - v.min()
    return v.reshape(mesh, mesh)


def projection_x(kind, temp, xrange, yrange, mesh=200, params=None, bias_fn=None):
    V = grid_energy(kind, xrange, yrange, mesh, params, bias_fn) / (KB_PY * temp)
    Fx = -_lse(-V, 0)
    return np.linspace(xrange[0], xrange[1], mesh), Fx - Fx.min()


def projection_y(kind, temp, xrange, yrange, mesh=200, params=None, bias_fn=None):
    V = grid_energy(kind, xrange, yrange, mesh, params, bias_fn) / (KB_PY * temp)
    Fy = -_lse(-V, 1)
    return np.linspace(yrange[0], yrange[1], mesh), Fy - Fy.min()


def counts_to_fes(counts):
    counts = np.array(counts, dtype=np.float64)
    counts[counts == 0] = counts[counts != 0].min()
    bF = -np.log(counts)
    return bF - bF.min()


def histogram_fes_2d(xvals, yvals, xrange, yrange, nbins_x=100, nbins_y=100, weights=None):
    counts, _, _ = np.histogram2d(xvals, yvals, bins=[nbins_x, nbins_y],
                                  range=[xrange, yrange], weights=weights)
    return counts_to_fes(counts)


def histogram_fes_1d(vals, vrange, nbins=100, weights=None):
    counts, edges = np.histogram(vals, bins=nbins, range=vrange, weights=weights)
    return 0.5 * (edges[:-1] + edges[1:]), counts_to_fes(counts)


def rmse_kt(F_est, F_ref, clip=20.0):
    """RMSE between two min-shifted profiles (kT) where F_ref <= clip."""
    m = np.asarray(F_ref) <= clip
    a = np.asarray(F_est)[m]
    b = np.asarray(F_ref)[m]
    a = a - a.min()
    b = b - b.min()
    return float(np.sqrt(np.mean((a - b) ** 2)))
