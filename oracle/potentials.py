"""Oracle (test infrastructure): analytic potential energy surfaces and their gradients.

Energies follow the numpy twins of the reference (``ves/potentials.py``: single wells
``:54-65,128-141``, double wells ``:67-77,144-157``, slip bond ``:160-189``, catch bond
``:192-233``, Szabo-Berezhkovskii ``:236-267``, Muller-Brown ``:270-297``), which are the same
formulas as the Lepton strings OpenMM evaluates.  The z restraint ``1000 z^2`` (``:105``) and,
for the 1D surfaces, ``1000 y^2`` (``:32``) are included.  Gradients are derived by hand
(the reference gets them from Lepton's symbolic differentiation inside OpenMM); tests check
them against central differences of the energies and the energies against golden values
produced by the reference's own ``.potential()``.

Kind ids and parameter vectors are the ones of ``include/pyves_b200.h``.
"""
import numpy as np

SINGLE_WELL_1D = 0
DOUBLE_WELL_1D = 1
SINGLE_WELL_2D = 2
DOUBLE_WELL_2D = 3
SLIP_BOND_2D = 4
CATCH_BOND_2D = 5
SZABO_BEREZHKOVSKII = 6
MULLER_BROWN = 7

SLIP_DEFAULT = (0.0, 0.0, 1.0, 5.0, 4.0, 0.0, 2.0)                     # fx fy y0 ys ysh xy0 xys
CATCH_DEFAULT = SLIP_DEFAULT + (2.0, 0.5, -2.5, 0.25)                  # + gx0 gxs gy0 gys
SB_DEFAULT = (2.2, 4.0, 1.01 * 4.0, 4.0 * 2.2 ** 2 / 4.0)              # x0 omega2 Omega2 Delta
MB_A = (-200.0, -100.0, -170.0, 15.0)
MB_a = (-1.0, -1.0, -6.5, 0.7)
MB_b = (0.0, 0.0, 11.0, 0.6)
MB_c = (-10.0, -10.0, -6.5, 0.7)
MB_X = (1.0, 0.0, -0.5, -1.0)
MB_Y = (0.0, 0.5, 1.5, 1.0)

K_RESTRAINT = 1000.0


def default_params(kind):
    if kind == SLIP_BOND_2D:
        return SLIP_DEFAULT
    if kind == CATCH_BOND_2D:
        return CATCH_DEFAULT
    if kind == SZABO_BEREZHKOVSKII:
        return SB_DEFAULT
    return ()


def energy_grad_xy(kind, x, y, params=None):
    """U(x, y) without the z restraint, and (dU/dx, dU/dy).  1D kinds include 1000 y^2."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    p = tuple(params) if params is not None and len(params) else default_params(kind)
    if kind == SINGLE_WELL_1D:
        return x ** 2 + K_RESTRAINT * y ** 2, 2 * x, 2 * K_RESTRAINT * y
    if kind == DOUBLE_WELL_1D:
        return (x ** 4 - 4 * x ** 2 + 0.7 * x + K_RESTRAINT * y ** 2,
                4 * x ** 3 - 8 * x + 0.7, 2 * K_RESTRAINT * y)
    if kind == SINGLE_WELL_2D:
        return x ** 2 + y ** 2, 2 * x, 2 * y
    if kind == DOUBLE_WELL_2D:
        return (10 * (x ** 4 - 4 * x ** 2 + 0.7 * x) + 5 * y ** 2,
                10 * (4 * x ** 3 - 8 * x + 0.7), 10 * y)
    if kind in (SLIP_BOND_2D, CATCH_BOND_2D):
        fx, fy, y0, ys, ysh, xy0, xys = p[:7]
        t = (y - y0) ** 2 / ys - ysh
        u = x - y - xy0
        A = t ** 2 + u ** 2 / xys
        dAx = 2 * u / xys
        dAy = 4 * t * (y - y0) / ys - 2 * u / xys
        if kind == SLIP_BOND_2D:
            return A - fx * x - fy * y, dAx - fx, dAy - fy
        gx0, gxs, gy0, gys = p[7:11]
        B = (x - gx0) ** 2 / gxs + (y - gy0) ** 2 / gys
        dBx = 2 * (x - gx0) / gxs
        dBy = 2 * (y - gy0) / gys
        m = np.minimum(A, B)
        ea = np.exp(-(A - m))
        eb = np.exp(-(B - m))
        U = m - np.log(ea + eb)
        wa = ea / (ea + eb)
        wb = eb / (ea + eb)
        return U - fx * x - fy * y, wa * dAx + wb * dBx - fx, wa * dAy + wb * dBy - fy
    if kind == SZABO_BEREZHKOVSKII:
        x0, w2, W2, D = p[:4]
        left = x < -0.5 * x0          # Lepton: step(x + x0/2) == 0
        right = x >= 0.5 * x0         # Lepton: step(x - x0/2) == 1
        Ux = np.where(left, -D + 0.5 * w2 * (x + x0) ** 2,
                      np.where(right, -D + 0.5 * w2 * (x - x0) ** 2, -0.5 * w2 * x ** 2))
        dUx = np.where(left, w2 * (x + x0), np.where(right, w2 * (x - x0), -w2 * x))
        return Ux + 0.5 * W2 * (x - y) ** 2, dUx + W2 * (x - y), -W2 * (x - y)
    if kind == MULLER_BROWN:
        U = np.zeros(np.broadcast(x, y).shape)
        gx = np.zeros_like(U)
        gy = np.zeros_like(U)
        for i in range(4):
            dx = x - MB_X[i]
            dy = y - MB_Y[i]
            e = MB_A[i] * np.exp(MB_a[i] * dx ** 2 + MB_b[i] * dx * dy + MB_c[i] * dy ** 2)
            U = U + e
            gx = gx + e * (2 * MB_a[i] * dx + MB_b[i] * dy)
            gy = gy + e * (MB_b[i] * dx + 2 * MB_c[i] * dy)
        return U, gx, gy
    raise ValueError("unknown potential kind %r" % (kind,))


def energy(kind, pos, params=None):
    """U of positions [N, 2 or 3] including the z restraint when a z column is present."""
    pos = np.asarray(pos, dtype=np.float64)
    U, _, _ = energy_grad_xy(kind, pos[:, 0], pos[:, 1], params)
    if pos.shape[1] > 2:
        U = U + K_RESTRAINT * pos[:, 2] ** 2
    return U


def force(kind, pos, params=None):
    """-grad U for positions [N, D], D = 2 or 3."""
    pos = np.asarray(pos, dtype=np.float64)
    _, gx, gy = energy_grad_xy(kind, pos[:, 0], pos[:, 1], params)
    cols = [-gx, -gy]
    if pos.shape[1] > 2:
        cols.append(-2 * K_RESTRAINT * pos[:, 2])
    return np.stack(cols, axis=1)
