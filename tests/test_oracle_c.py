"""The plain-C oracle port (CPU baseline of bench.py) against the numpy oracle."""
import numpy as np

from oracle import cport, langevin, mlp, potentials
from tests import biasspec


def _case(pot_kind, params, spec, box, seed=3, n=64, nsteps=41):
    rng = np.random.default_rng(seed)
    k = langevin.constants()
    x0 = np.stack([rng.uniform(box[0], box[1], n), rng.uniform(box[2], box[3], n)], 1)
    v0 = rng.standard_normal((n, 2)) * k["sigma_v"]
    f = biasspec.total_force_fn(pot_kind, params, spec)
    xr, vr = langevin.run_philox(x0, v0, f, k, seed, 1000, 5, nsteps)
    m = cport.Model(pot_kind, params or potentials.default_params(pot_kind), spec, k, seed=seed, gid0=1000)
    pos = np.ascontiguousarray(x0.T)
    vel = np.ascontiguousarray(v0.T)
    m.run(pos, vel, 5, nsteps)
    assert np.allclose(pos.T, xr, rtol=1e-10, atol=1e-10) and np.allclose(vel.T, vr, rtol=1e-10, atol=1e-10)
    return m, pos


def test_c_port_matches_numpy_oracle():
    rng = np.random.default_rng(0)
    leg2 = dict(kind="leg2d", deg_x=15, deg_y=15, minmax=(-2.5, 2.5, -2.0, 2.0), weights=rng.standard_normal((16, 16)) * 0.5)
    m, pos = _case(potentials.DOUBLE_WELL_2D, None, leg2, (-1.6, -1.2, -0.3, 0.3))
    from oracle import legendre
    assert np.allclose(m.moments(pos), legendre.basis2d(pos.T, 15, 15, -2.5, 2.5, -2.0, 2.0).sum(0))
    leg1 = dict(kind="leg1d", axis='x', degree=12, min=-3.0, max=6.0, weights=rng.standard_normal(13))
    m, pos = _case(potentials.SLIP_BOND_2D, None, leg1, (-3.7, -3.2, -3.7, -3.2))
    assert np.allclose(m.moments(pos), legendre.basis1d(pos.T, 12, -3.0, 6.0, 'x').sum(0))
    nn = dict(kind="mlp", axis='x', min=-4.0, max=4.0, sizes=[48, 48, 48], theta=rng.standard_normal(mlp.n_params([48, 48, 48])) * 0.3)
    _case(potentials.SZABO_BEREZHKOVSKII, None, nn, (2.0, 2.4, 2.0, 2.4))
    _case(potentials.DOUBLE_WELL_2D, None, dict(kind="none"), (-1.6, -1.2, -0.3, 0.3))
    assert cport.num_threads() >= 1
