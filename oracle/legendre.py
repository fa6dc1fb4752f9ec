"""Oracle (test infrastructure): Legendre basis expansions.

1D follows ``ves/basis.py:123-156`` (``LegendreBasis1D.forward``: axis pick ``:135-140``,
scaling ``:143``, clamp ``:147``) and ``ves/basis.py:86-121`` (Bonnet recursion
``P_{n+1} = ((2n+1) x P_n - n P_{n-1}) / (n+1)``, ``out = sum_n w_n P_n``).  The reference
obtains forces by autograd; torch's clamp passes the gradient where ``-1 <= s <= 1``
inclusive and blocks it outside, which is restated here as the ``inside`` mask.

The 2D tensor-product basis has no reference symbol: it applies the same
per-axis scaling / clamp and ``V = sum_ij w_ij P_i(x') P_j(y')``.
"""
import numpy as np


def legendre_table(s, degree):
    """P_n(s) and dP_n/ds for n = 0..degree -> two arrays [degree+1, len(s)]."""
    s = np.asarray(s, dtype=np.float64)
    P = np.zeros((degree + 1,) + s.shape)
    dP = np.zeros_like(P)
    P[0] = 1.0
    if degree >= 1:
        P[1] = s
        dP[1] = 1.0
    for n in range(1, degree):
        P[n + 1] = ((2 * n + 1) * s * P[n] - n * P[n - 1]) / (n + 1)
        dP[n + 1] = dP[n - 1] + (2 * n + 1) * P[n]
    return P, dP


def scale_clamp(q, lo, hi):
    """(q - (lo+hi)/2) / ((hi-lo)/2), clamped to [-1, 1]; mask of rows where the clamp is inactive."""
    s = (np.asarray(q, dtype=np.float64) - (lo + hi) / 2) / ((hi - lo) / 2)
    inside = (s >= -1.0) & (s <= 1.0)
    return np.clip(s, -1.0, 1.0), inside


def legendre1d(pos, degree, lo, hi, axis, weights):
    """Energy [N] and force [N, D] of LegendreBasis1D(degree, lo, hi, axis, weights)."""
    pos = np.asarray(pos, dtype=np.float64)
    a = {'x': 0, 'y': 1}[axis]
    w = np.asarray(weights, dtype=np.float64)
    s, inside = scale_clamp(pos[:, a], lo, hi)
    P, dP = legendre_table(s, degree)
    E = np.tensordot(w, P, axes=(0, 0))
    dEds = np.tensordot(w, dP, axes=(0, 0))
    F = np.zeros_like(pos)
    F[:, a] = -dEds * inside / ((hi - lo) / 2)
    return E, F


def basis1d(pos, degree, lo, hi, axis):
    """Basis-function values f_k = P_k(s) -> [N, degree+1] (the gradient of V w.r.t. the weights)."""
    pos = np.asarray(pos, dtype=np.float64)
    a = {'x': 0, 'y': 1}[axis]
    s, _ = scale_clamp(pos[:, a], lo, hi)
    P, _ = legendre_table(s, degree)
    return P.T.copy()


def legendre2d(pos, deg_x, deg_y, xlo, xhi, ylo, yhi, weights):
    """Energy [N] and force [N, D] of the tensor-product basis; weights [deg_x+1, deg_y+1]."""
    pos = np.asarray(pos, dtype=np.float64)
    w = np.asarray(weights, dtype=np.float64).reshape(deg_x + 1, deg_y + 1)
    sx, inx = scale_clamp(pos[:, 0], xlo, xhi)
    sy, iny = scale_clamp(pos[:, 1], ylo, yhi)
    Px, dPx = legendre_table(sx, deg_x)
    Py, dPy = legendre_table(sy, deg_y)
    A = w @ Py            # [Kx, N]
    B = w @ dPy
    E = np.sum(Px * A, axis=0)
    gx = np.sum(dPx * A, axis=0)
    gy = np.sum(Px * B, axis=0)
    F = np.zeros_like(pos)
    F[:, 0] = -gx * inx / ((xhi - xlo) / 2)
    F[:, 1] = -gy * iny / ((yhi - ylo) / 2)
    return E, F


def basis2d(pos, deg_x, deg_y, xlo, xhi, ylo, yhi):
    """f_{ij} = P_i(x') P_j(y') flattened row-major -> [N, (deg_x+1)(deg_y+1)]."""
    pos = np.asarray(pos, dtype=np.float64)
    sx, _ = scale_clamp(pos[:, 0], xlo, xhi)
    sy, _ = scale_clamp(pos[:, 1], ylo, yhi)
    Px, _ = legendre_table(sx, deg_x)
    Py, _ = legendre_table(sy, deg_y)
    return np.einsum('in,jn->nij', Px, Py).reshape(pos.shape[0], -1)
