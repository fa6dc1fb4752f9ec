"""Analytic potential energy surfaces -- host-side mirror of ``ves/potentials.py``.

Same class names, constructor keywords, ``.force`` expression strings and ``.potential()`` numpy
methods as the reference (``ves/potentials.py:12-297``).  Instead of subclassing
``openmm.CustomExternalForce`` (Lepton-compiled at run time), every class carries a ``kind`` id and
a ``params`` tuple that select the corresponding device functor of the CUDA library
(``csrc/potentials.cuh``); no expression is parsed at run time.
"""
import numpy as np

from . import _lib

_Z_RESTRAINT = " + 1000 * z^2"
_YZ_RESTRAINT = " + 1000 * y^2 + 1000 * z^2"


class _Potential:
    kind = None
    params = ()
    quiet = False

    def _finish(self, tail):
        self.force += tail
        if not _Potential.quiet:      # the reference prints the expression (ves/potentials.py:35,108)
            print("[Potential] Initializing potential with expression:\n" + self.force)

    # openmm.CustomExternalForce API the reference driver calls (ves/langevin_dynamics.py:46)
    def addParticle(self, index, parameters=()):
        return 0


class Potential1D(_Potential):
    """Base of the 1D surfaces; y and z feel ``1000 q^2`` (``ves/potentials.py:12-51``)."""

    def __init__(self):
        self._finish(_YZ_RESTRAINT)

    def potential(self, x):
        raise NotImplementedError()


class SingleWellPotential1D(Potential1D):
    kind = _lib.POT_SINGLE_WELL_1D

    def __init__(self):
        self.force = "x^2"
        super().__init__()

    def potential(self, x):
        return np.square(x)


class DoubleWellPotential1D(Potential1D):
    kind = _lib.POT_DOUBLE_WELL_1D

    def __init__(self):
        self.force = "x^4 - 4 * x^2 + 0.7 * x"
        super().__init__()

    def potential(self, x):
        x2 = np.square(x)
        return x2 * x2 - 4 * x2 + 0.7 * x


class Potential2D(_Potential):
    """Base of the 2D surfaces; z feels ``1000 z^2`` (``ves/potentials.py:85-125``)."""

    def __init__(self):
        self._finish(_Z_RESTRAINT)

    def potential(self, x, y):
        raise NotImplementedError()


class SingleWellPotential2D(Potential2D):
    kind = _lib.POT_SINGLE_WELL_2D

    def __init__(self):
        self.force = "x^2 + y^2"
        super().__init__()

    def potential(self, x, y):
        return np.square(x) + np.square(y)


class DoubleWellPotential2D(Potential2D):
    r"""$U = 10 (x^4 - 4 x^2 + 0.7 x) + 5 y^2$ (``ves/potentials.py:144-157``)."""
    kind = _lib.POT_DOUBLE_WELL_2D

    def __init__(self):
        self.force = "10 * (x^4 - 4 * x^2 + 0.7 * x) + 5 * y^2"
        super().__init__()

    def potential(self, x, y):
        x2 = np.square(x)
        return 10 * (x2 * x2 - 4 * x2 + 0.7 * x) + 5 * np.square(y)


_SLIP_KEYS = ("force_x", "force_y", "y_0", "y_scale", "y_shift", "xy_0", "xy_scale")
_CATCH_KEYS = _SLIP_KEYS + ("gx_0", "gx_scale", "gy_0", "gy_scale")
_SLIP_EXPR = "((y - {y_0})^2 / {y_scale} - {y_shift})^2 + (x - y - {xy_0})^2 / {xy_scale}"


class SlipBondPotential2D(Potential2D):
    """Two-basin slip bond surface (``ves/potentials.py:160-189``)."""
    kind = _lib.POT_SLIP_BOND_2D

    def __init__(self, force_x=0, force_y=0, y_0=1, y_scale=5, y_shift=4, xy_0=0, xy_scale=2):
        vals = dict(zip(_SLIP_KEYS, (force_x, force_y, y_0, y_scale, y_shift, xy_0, xy_scale)))
        self.__dict__.update(vals)
        self.params = tuple(float(vals[k]) for k in _SLIP_KEYS)
        self.force = (_SLIP_EXPR + " - {force_x} * x - {force_y} * y").format(**vals)
        super().__init__()

    def _basin(self, x, y):
        return (np.square(y - self.y_0) / self.y_scale - self.y_shift) ** 2 + np.square(x - y - self.xy_0) / self.xy_scale

    def potential(self, x, y):
        return self._basin(x, y) - self.force_x * x - self.force_y * y


class CatchBondPotential2D(SlipBondPotential2D):
    """Slip bond surface with an extra harmonic basin, combined by a soft minimum
    (``ves/potentials.py:192-233``)."""
    kind = _lib.POT_CATCH_BOND_2D

    def __init__(self, force_x=0, force_y=0, y_0=1, y_scale=5, y_shift=4, xy_0=0, xy_scale=2,
                 gx_0=2, gx_scale=0.5, gy_0=-2.5, gy_scale=0.25):
        vals = dict(zip(_CATCH_KEYS, (force_x, force_y, y_0, y_scale, y_shift, xy_0, xy_scale,
                                      gx_0, gx_scale, gy_0, gy_scale)))
        self.__dict__.update(vals)
        self.params = tuple(float(vals[k]) for k in _CATCH_KEYS)
        self.force = ("-log(exp(-(" + _SLIP_EXPR + ") ) + exp(-( (x - {gx_0})^2 / {gx_scale} + (y - {gy_0})^2 / {gy_scale} )) )"
                      " - {force_x} * x - {force_y} * y").format(**vals)
        Potential2D.__init__(self)

    def potential(self, x, y):
        well = np.square(x - self.gx_0) / self.gx_scale + np.square(y - self.gy_0) / self.gy_scale
        return -np.log(np.exp(-self._basin(x, y)) + np.exp(-well)) - self.force_x * x - self.force_y * y


class SzaboBerezhkovskiiPotential(Potential2D):
    """Szabo-Berezhkovskii surface: harmonic coupling of y to a piecewise-parabolic double well in x
    (``ves/potentials.py:236-267``)."""
    kind = _lib.POT_SZABO_BEREZHKOVSKII
    x0 = 2.2
    omega2 = 4.0
    Omega2 = 1.01 * omega2
    Delta = omega2 * x0 ** 2 / 4.0

    def __init__(self):
        c = dict(x0=self.x0, omega2=self.omega2, Omega2=self.Omega2, Delta=self.Delta)
        self.params = (self.x0, self.omega2, self.Omega2, self.Delta)
        outer = "-{Delta} + {omega2} * 0.5 * (x {sign} {x0})^2"
        self.force = ("{Omega2} * 0.5 * (x - y)^2 + (select(step(x + 0.5 * {x0}), select(step(x - 0.5 * {x0}), "
                      + outer.replace("{sign}", "-") + ", -{omega2} * 0.5 * x^2), " + outer.replace("{sign}", "+") + "))").format(**c)
        super().__init__()

    def potential(self, x, y):
        x = np.asarray(x, dtype=float)
        half = self.x0 / 2
        wells = -self.Delta + self.omega2 * np.square(np.abs(x) - self.x0) / 2.0
        Ux = np.where(np.abs(x) >= half, wells, -self.omega2 * np.square(x) / 2.0)
        # the reference's np.piecewise puts |x| == x0/2 in the outer wells; the pieces agree there
        return Ux + self.Omega2 * np.square(x - y) / 2.0


class MullerBrownPotential(Potential2D):
    """Muller-Brown surface, four Gaussians (``ves/potentials.py:270-297``)."""
    kind = _lib.POT_MULLER_BROWN
    a = [-1, -1, -6.5, 0.7]
    b = [0, 0, 11, 0.6]
    c = [-10, -10, -6.5, 0.7]
    A = [-200, -100, -170, 15]
    x_bar = [1, 0, -0.5, -1]
    y_bar = [0, 0.5, 1.5, 1]

    def __init__(self):
        term = "{A} * exp({a} * (x - {x_bar})^2 + {b} * (x - {x_bar}) * (y - {y_bar}) + {c} * (y - {y_bar})^2)"
        self.force = " + ".join(term.format(a=self.a[i], b=self.b[i], c=self.c[i], A=self.A[i],
                                            x_bar=self.x_bar[i], y_bar=self.y_bar[i]) for i in range(4))
        super().__init__()

    def potential(self, x, y):
        dx = np.subtract.outer(np.asarray(x, dtype=float), self.x_bar)
        dy = np.subtract.outer(np.asarray(y, dtype=float), self.y_bar)
        return np.sum(np.asarray(self.A) * np.exp(np.asarray(self.a) * dx * dx + np.asarray(self.b) * dx * dy
                                                   + np.asarray(self.c) * dy * dy), axis=-1)
