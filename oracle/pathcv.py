"""Oracle (test infrastructure): Legendre expansion along a path collective variable.

Follows ``ves/basis.py:509-548`` (``LegendreBasis2DPathCV.forward``):
``d_i = (x-x_i)^2 + (y-y_i)^2``; ``s = (1/N) exp(LSE_i(-lam d_i + ln(i+1)) - LSE_i(-lam d_i))``
(``:524-526``); ``z = -(1/lam) LSE_i(-lam d_i)`` (``:529``); ``s' = clamp(2 s - 1, -1, 1)``
(``:532-535``); ``V = sum_n w_n P_n(s') + phi z`` (``:543-546``).  Forces by hand-derived
chain rule (the reference uses autograd).  Also the harmonic bias of ``ves/bias.py:133``.
"""
import numpy as np

from .legendre import legendre_table


def _lse(a, axis):
    m = a.max(axis=axis, keepdims=True)
    return (m + np.log(np.exp(a - m).sum(axis=axis, keepdims=True))).squeeze(axis)


def path_cv(pos, x_i, y_i, lam):
    """s, z and their gradients w.r.t. (x, y): returns s, z, ds[N,2], dz[N,2]."""
    pos = np.asarray(pos, dtype=np.float64)
    x_i = np.asarray(x_i, dtype=np.float64)
    y_i = np.asarray(y_i, dtype=np.float64)
    N = len(x_i)
    dx = pos[:, 0][:, None] - x_i[None, :]
    dy = pos[:, 1][:, None] - y_i[None, :]
    a = -lam * (dx ** 2 + dy ** 2)
    lni = np.log(np.arange(1, N + 1, dtype=np.float64))
    L0 = _lse(a, 1)
    L1 = _lse(a + lni[None, :], 1)
    p = np.exp(a - L0[:, None])
    q = np.exp(a + lni[None, :] - L1[:, None])
    dL0 = np.stack([(p * (-2 * lam * dx)).sum(1), (p * (-2 * lam * dy)).sum(1)], axis=1)
    dL1 = np.stack([(q * (-2 * lam * dx)).sum(1), (q * (-2 * lam * dy)).sum(1)], axis=1)
    s = np.exp(L1 - L0) / N
    z = -L0 / lam
    return s, z, s[:, None] * (dL1 - dL0), -dL0 / lam


def legendre_pathcv(pos, degree, x_i, y_i, lam, weights, phi=0.0):
    """Energy [N] and force [N, D]."""
    pos = np.asarray(pos, dtype=np.float64)
    w = np.asarray(weights, dtype=np.float64)
    s, z, ds, dz = path_cv(pos, x_i, y_i, lam)
    sp = (s - 0.5) / 0.5
    inside = (sp >= -1.0) & (sp <= 1.0)
    sp = np.clip(sp, -1.0, 1.0)
    P, dP = legendre_table(sp, degree)
    E = np.tensordot(w, P, axes=(0, 0)) + phi * z
    dEds = np.tensordot(w, dP, axes=(0, 0)) * inside * 2.0
    F = np.zeros_like(pos)
    F[:, :2] = -(dEds[:, None] * ds + phi * dz)
    return E, F


def harmonic(pos, k, s0, axis):
    """Per-row k/2 (q - s0)^2 and its force (``ves/bias.py:133`` sums over rows; N = 1 there)."""
    pos = np.asarray(pos, dtype=np.float64)
    a = {'x': 0, 'y': 1}[axis]
    E = 0.5 * k * (pos[:, a] - s0) ** 2
    F = np.zeros_like(pos)
    F[:, a] = -k * (pos[:, a] - s0)
    return E, F
