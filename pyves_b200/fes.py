"""Free-energy estimates -- the numerical content of ``ves/visualization.py`` without plotting.

* ``projection_x`` / ``projection_y``: analytic F along x / y of a ``Potential2D`` on a mesh
  (``ves/visualization.py:185-201,221-237``; host numpy, a 200 x 200 grid evaluated once).
* ``histogram_fes_1d`` / ``histogram_fes_2d``: ``beta F = -log counts`` with empty bins set to the
  smallest non-zero count (``ves/visualization.py:475-478``), the counts coming from the CUDA
  histogram kernel (``pyves_histogram``) over an ensemble or over caller positions.
"""
import numpy as np

KB_PY = 8.3145 / 1000


def _lse(a, axis):
    m = a.max(axis=axis, keepdims=True)
    return (m + np.log(np.exp(a - m).sum(axis=axis, keepdims=True))).squeeze(axis)


def _grid(potential2D, temp, xrange, yrange, mesh):
    xx, yy = np.meshgrid(np.linspace(xrange[0], xrange[1], mesh), np.linspace(yrange[0], yrange[1], mesh))
    return potential2D.potential(xx.ravel(), yy.ravel()).reshape(mesh, mesh) / (KB_PY * temp)


def projection_x(potential2D, temp, xrange, yrange, mesh=200):
    """(x [mesh], F_x [mesh] in kT, min-shifted)"""
    F = -_lse(-_grid(potential2D, temp, xrange, yrange, mesh), 0)
    return np.linspace(xrange[0], xrange[1], mesh), F - F.min()


def projection_y(potential2D, temp, xrange, yrange, mesh=200):
    F = -_lse(-_grid(potential2D, temp, xrange, yrange, mesh), 1)
    return np.linspace(yrange[0], yrange[1], mesh), F - F.min()


def counts_to_fes(counts):
    counts = np.array(counts, dtype=np.float64)
    if not np.any(counts > 0):
        raise ValueError("empty histogram")
    counts[counts == 0] = counts[counts != 0].min()
    bF = -np.log(counts)
    return bF - bF.min()


def histogram_fes_1d(ens, vrange, nbins=100, axis=0, pos=None, row_weights=None, counts=None):
    """(bin centres, beta F) from the walkers of ``ens`` (or rows ``pos``); pass ``counts`` to accumulate."""
    counts = ens.histogram(vrange, nbins, axis=axis, pos=pos, row_weights=row_weights, counts=counts)
    c = counts.cpu().numpy() if hasattr(counts, "cpu") else counts
    edges = np.linspace(vrange[0], vrange[1], nbins + 1)
    return 0.5 * (edges[:-1] + edges[1:]), counts_to_fes(c)


def histogram_fes_2d(ens, xrange, yrange, nbins_x=100, nbins_y=100, pos=None, row_weights=None, counts=None):
    counts = ens.histogram((xrange[0], xrange[1], yrange[0], yrange[1]), (nbins_x, nbins_y), pos=pos,
                           row_weights=row_weights, counts=counts)
    c = counts.cpu().numpy() if hasattr(counts, "cpu") else counts
    return counts_to_fes(c)


def rmse_kt(F_est, F_ref, clip=20.0):
    """RMSE in kT between two min-shifted profiles where the reference is <= clip."""
    m = np.asarray(F_ref) <= clip
    a, b = np.asarray(F_est)[m], np.asarray(F_ref)[m]
    return float(np.sqrt(np.mean(((a - a.min()) - (b - b.min())) ** 2)))
