"""Target distributions p(s) -- mirror of ``ves/target.py`` plus the well-tempered target.

``Target_x`` / ``Target_Uniform_HardSwitch_x`` keep the reference contract (``ves/target.py:7-31``):
``.x`` nodes and ``.p`` densities as numpy arrays; the uniform density is 1/2 at every node of
``linspace(-1, 1, mesh)`` (a density, not normalised weights).

New (named by the reference only in its README flow-chart, ``README.md:10-14``): the well-tempered
target p(s) ~ exp(-(beta/gamma) F(s)) (Valsson & Parrinello, JCTC 11, 1996 (2015)), iterated from the
current bias as F = -V - (1/beta) ln p, in 1D and on a 2D tensor grid.  Nodes live in the scaled
coordinate s in [-1, 1] of the basis the target is paired with.
"""
import numpy as np


class Target_x:
    """Abstract target distribution along one collective variable."""

    def __init__(self):
        self._x = None
        self._p = None

    @property
    def x(self):
        return self._x

    @property
    def p(self):
        return self._p

    # ---- additions used by the ensemble optimiser ----
    ndim = 1

    def quadrature_weights(self):
        """Normalised node weights for <f>_p = sum_i q_i f(x_i) (trapezoid rule on the mesh)."""
        q = np.array(self._p, dtype=np.float64)
        if q.size > 1:
            q[0] *= 0.5
            q[-1] *= 0.5
        return q / q.sum()

    def update(self, V_nodes, beta):
        """Refresh the target from the bias evaluated at the nodes; static targets do nothing."""
        return False


class Target_Uniform_HardSwitch_x(Target_x):
    """Uniform over [-1, 1], zero outside (``ves/target.py:24-31``)."""

    def __init__(self, mesh):
        super().__init__()
        self._x = np.linspace(-1, 1, mesh)
        self._p = np.full(self._x.shape, 0.5)


class Target_WellTempered_x(Target_x):
    """Well-tempered target on ``linspace(-1, 1, mesh)``: starts uniform, ``update`` re-derives it
    from the bias, p' ~ exp(-(beta / bias_factor) F), F = -V - (1/beta) ln p."""

    def __init__(self, mesh, bias_factor):
        super().__init__()
        if not bias_factor > 1:
            raise ValueError("bias_factor must be > 1")
        self.bias_factor = float(bias_factor)
        self._x = np.linspace(-1, 1, mesh)
        self._p = np.full(self._x.shape, 0.5)

    def update(self, V_nodes, beta):
        F = -np.asarray(V_nodes, dtype=np.float64) - np.log(np.maximum(self._p, 1e-300)) / beta
        q = np.exp(-(beta / self.bias_factor) * (F - F.min()))
        w = np.ones_like(q)
        w[0] = w[-1] = 0.5
        self._p = q / (np.sum(q * w) * (self._x[1] - self._x[0]))     # density: integrates to 1 on [-1, 1]
        return True


class Target_2D:
    """Target on a tensor grid of cell centres in scaled coordinates [-1, 1]^2.

    ``.x`` [M, 2] nodes (x fastest varying last: row-major over (ix, iy)), ``.p`` [M] densities."""
    ndim = 2

    def __init__(self, mesh_x, mesh_y):
        self.mesh = (int(mesh_x), int(mesh_y))
        cx = -1 + (2 * np.arange(self.mesh[0]) + 1) / self.mesh[0]
        cy = -1 + (2 * np.arange(self.mesh[1]) + 1) / self.mesh[1]
        gx, gy = np.meshgrid(cx, cy, indexing="ij")
        self._x = np.stack([gx.ravel(), gy.ravel()], axis=1)
        self._p = np.full(self._x.shape[0], 0.25)

    @property
    def x(self):
        return self._x

    @property
    def p(self):
        return self._p

    def quadrature_weights(self):
        return self._p / self._p.sum()      # midpoint rule, equal cells

    def update(self, V_nodes, beta):
        return False


class Target_Uniform_2D(Target_2D):
    """Uniform over [-1, 1]^2."""


class Target_WellTempered_2D(Target_2D):
    """Well-tempered target on the 2D grid (BASELINE.json configs[1])."""

    def __init__(self, mesh_x, mesh_y, bias_factor):
        super().__init__(mesh_x, mesh_y)
        if not bias_factor > 1:
            raise ValueError("bias_factor must be > 1")
        self.bias_factor = float(bias_factor)

    def update(self, V_nodes, beta):
        F = -np.asarray(V_nodes, dtype=np.float64) - np.log(np.maximum(self._p, 1e-300)) / beta
        q = np.exp(-(beta / self.bias_factor) * (F - F.min()))
        cell = 4.0 / (self.mesh[0] * self.mesh[1])
        self._p = q / (q.sum() * cell)
        return True
