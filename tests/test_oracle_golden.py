"""The oracle against the fixtures produced by the unmodified reference (tests/golden/make_golden.py)
and against the known answers the reference's own tests use (tests/tests_unit/test_basis.py)."""
import numpy as np
import pytest
from scipy.special import legendre as scipy_legendre

from oracle import fes, langevin, legendre, mlp, pathcv, philox, potentials, ves

RTOL = 1e-10


def test_legendre_polynomials_vs_scipy():
    # reference test_basis.py:11-25 (values) and :48-68 (derivatives), extended to degree 12
    x = np.linspace(-1, 1, 1000)
    P, dP = legendre.legendre_table(x, 12)
    for n in range(13):
        assert np.allclose(P[n], scipy_legendre(n)(x), rtol=1e-9, atol=1e-11)
        assert np.allclose(dP[n], scipy_legendre(n).deriv()(x), rtol=1e-9, atol=1e-9)


def test_legendre_expansion_vs_single_polynomials():
    # reference test_basis.py:28-45
    x = np.linspace(-1, 1, 1000)
    w = np.random.default_rng(0).random(6)
    P, _ = legendre.legendre_table(x, 5)
    pos = np.stack([x, 0 * x, 0 * x], 1)
    E, _ = legendre.legendre1d(pos, 5, -1.0, 1.0, 'x', w)
    assert np.allclose(E, sum(w[i] * P[i] for i in range(6)))


def test_legendre1d_golden(golden):
    for c in golden("legendre1d.json"):
        E, F = legendre.legendre1d(c["positions"], c["degree"], c["min"], c["max"], c["axis"], c["weights"])
        assert np.allclose(E, c["E"], rtol=RTOL, atol=1e-11), c["tag"]
        assert np.allclose(F, c["F"], rtol=RTOL, atol=1e-9), c["tag"]


def test_legendre1d_survey_appendix_b1():
    # clamp is inclusive at the ends, flat outside
    w = [-4.48286631e+00, -1.74999996e+00, -5.65477074e+00, 7.26710048e-05, -8.92828224e+00]
    pos = np.array([[-3.0, 0.7, 0], [-2.5, 0.7, 0], [2.5, 0.7, 0], [2.6, 0.7, 0]])
    E, F = legendre.legendre1d(pos, 4, -2.5, 2.5, 'x', w)
    assert np.allclose(E, [-17.315992001005, -17.315992001005, -20.815846578995, -20.815846578995])
    assert np.allclose(F[:, 0], [0, -41.799028274412, 43.198679421588, 0])


def test_legendre2d_vs_numpy_legval2d():
    rng = np.random.default_rng(1)
    w = rng.standard_normal((16, 16))
    pos = np.stack([rng.uniform(-3, 3, 200), rng.uniform(-2.5, 2.5, 200)], 1)
    E, F = legendre.legendre2d(pos, 15, 15, -2.5, 2.5, -2.0, 2.0, w)
    sx = np.clip(pos[:, 0] / 2.5, -1, 1)
    sy = np.clip(pos[:, 1] / 2.0, -1, 1)
    assert np.allclose(E, np.polynomial.legendre.legval2d(sx, sy, w), rtol=1e-10, atol=1e-10)
    # forces vs central differences (inside the box only)
    h = 1e-6
    for a in range(2):
        d = np.zeros_like(pos)
        d[:, a] = h
        Ep, _ = legendre.legendre2d(pos + d, 15, 15, -2.5, 2.5, -2.0, 2.0, w)
        Em, _ = legendre.legendre2d(pos - d, 15, 15, -2.5, 2.5, -2.0, 2.0, w)
        lim = (2.5, 2.0)[a]
        m = np.abs(pos[:, a]) < lim - 1e-3
        assert np.allclose(F[m, a], -(Ep - Em)[m] / (2 * h), rtol=1e-5, atol=1e-4)
        assert np.all(F[np.abs(pos[:, a]) > lim, a] == 0)
    f = legendre.basis2d(pos, 15, 15, -2.5, 2.5, -2.0, 2.0)
    assert np.allclose(f @ w.ravel(), E)


def test_nnbasis_golden(golden):
    for c in golden("nnbasis1d.json"):
        act = {"relu": mlp.ACT_RELU, "tanh": mlp.ACT_TANH}[c["act"]]
        E, F = mlp.nnbasis1d(c["positions"], c["min"], c["max"], c["axis"], c["theta"], c["sizes"], act)
        assert len(c["theta"]) == mlp.n_params(c["sizes"])
        assert np.allclose(E, c["E"], rtol=RTOL, atol=1e-11), c["tag"]
        assert np.allclose(F, c["F"], rtol=RTOL, atol=1e-10), c["tag"]


def test_pathcv_and_harmonic_golden(golden):
    for c in golden("pathcv.json"):
        E, F = pathcv.legendre_pathcv(c["positions"], c["degree"], c["x_i"], c["y_i"], c["lam"], c["weights"], c["phi"])
        assert np.allclose(E, c["E"], rtol=1e-9, atol=1e-10), c["tag"]
        assert np.allclose(F, c["F"], rtol=1e-8, atol=1e-8), c["tag"]
    c = golden("harmonic.json")
    E, F = pathcv.harmonic(c["positions"], c["k"], c["s0"], c["axis"])
    assert np.allclose(E, c["E"]) and np.allclose(F, c["F"])


NAME_TO_KIND = dict(SingleWellPotential2D=potentials.SINGLE_WELL_2D, DoubleWellPotential2D=potentials.DOUBLE_WELL_2D,
                    SlipBondPotential2D=potentials.SLIP_BOND_2D, SlipBondPotential2D_forced=potentials.SLIP_BOND_2D,
                    CatchBondPotential2D=potentials.CATCH_BOND_2D,
                    SzaboBerezhkovskiiPotential=potentials.SZABO_BEREZHKOVSKII,
                    MullerBrownPotential=potentials.MULLER_BROWN,
                    SingleWellPotential1D=potentials.SINGLE_WELL_1D, DoubleWellPotential1D=potentials.DOUBLE_WELL_1D)
FORCED = (1.5, -0.5, 0.8, 4.0, 3.0, 0.2, 2.5)


def test_potentials_golden_and_gradients(golden):
    g = golden("potentials.json")
    for name, c in g.items():
        kind = NAME_TO_KIND[name]
        params = FORCED if name.endswith("_forced") else None
        x = np.array(c["x"])
        y = np.array(c["y"]) if "y" in c else np.zeros_like(x)
        U, gx, gy = potentials.energy_grad_xy(kind, x, y, params)
        assert np.allclose(U, c["U"], rtol=1e-12, atol=1e-12), name
        h = 1e-6
        Up, _, _ = potentials.energy_grad_xy(kind, x + h, y, params)
        Um, _, _ = potentials.energy_grad_xy(kind, x - h, y, params)
        assert np.allclose(gx, (Up - Um) / (2 * h), rtol=2e-5, atol=2e-4), name
        Up, _, _ = potentials.energy_grad_xy(kind, x, y + h, params)
        Um, _, _ = potentials.energy_grad_xy(kind, x, y - h, params)
        assert np.allclose(gy, (Up - Um) / (2 * h), rtol=2e-5, atol=2e-4), name


def test_update_literal_loss_golden(golden):
    """ves/bias.py:206-253 -- printed loss, parameter gradient and the optimiser step."""
    for c in golden("update.json"):
        traj = np.array(c["traj"])
        beta = c["beta"]
        xt = np.linspace(-1, 1, c["mesh"])
        pt = 0.5 * np.ones(c["mesh"])
        tgt_pos = np.stack([xt, 0 * xt, 0 * xt], 1)
        theta = np.array(c["theta0"])
        state = None
        for u in range(c["nupdates"]):
            tr = traj[: len(traj) - (c["nupdates"] - 1 - u) * 20]
            lo, hi = ves.window(len(tr), c["update_steps"])
            win = tr[lo:hi]
            if c["tag"].startswith("nn"):
                sizes = [8, 8]
                s_tr = (win[:, 0] + 4.0) / 8.0
                s_tg = (xt + 4.0) / 8.0
                V_tr, _ = mlp.forward(s_tr, theta, sizes)
                V_tg, _ = mlp.forward(s_tg, theta, sizes)
                dV_tr = np.stack([mlp.param_grad(s_tr[i:i + 1], theta, sizes) for i in range(len(s_tr))])
                dV_tg = np.stack([mlp.param_grad(s_tg[i:i + 1], theta, sizes) for i in range(len(s_tg))])
            else:
                V_tr, _ = legendre.legendre1d(win, 5, -2.5, 2.5, 'x', theta)
                V_tg, _ = legendre.legendre1d(tgt_pos, 5, -2.5, 2.5, 'x', theta)
                dV_tr = legendre.basis1d(win, 5, -2.5, 2.5, 'x')
                dV_tg = legendre.basis1d(tgt_pos, 5, -2.5, 2.5, 'x')
            loss, grad = ves.literal_loss_and_grad(V_tr, dV_tr, V_tg, dV_tg, pt, beta)
            assert abs(loss - c["loss_printed"][u]) < 1e-6, c["tag"]
            assert np.allclose(grad, c["grad"][u], rtol=1e-8, atol=1e-10), c["tag"]
            if c["optimizer"] == "SGD":
                theta, state = ves.sgd_step(theta, grad, **c["optimizer_params"], state=state)
            elif c["optimizer"] == "Adam":
                theta, state = ves.adam_step(theta, grad, **c["optimizer_params"], state=state)
            else:
                theta, state = ves.rmsprop_step(theta, grad, **c["optimizer_params"], state=state)
            assert np.allclose(theta, c["theta"][u], rtol=1e-8, atol=1e-10), c["tag"]


def test_update_b4_known_answer(golden):
    c = [x for x in golden("update.json") if x["tag"] == "B4_sgd"][0]
    assert abs(c["loss_printed"][0] - 11.486906) < 1e-6
    assert np.allclose(c["grad"][0], [6.5, -0.56202282, -2.24600345, 0.39300457, 0.55181266, 0.05861705], atol=1e-7)
    assert c["files"] == ["model.pt", "model.pt.iter200"]


def test_fes_projection_golden(golden):
    for name, c in golden("fes.json").items():
        kind = NAME_TO_KIND[name]
        x, Fx = fes.projection_x(kind, c["temp"], c["xrange"], c["yrange"], c["mesh"])
        y, Fy = fes.projection_y(kind, c["temp"], c["xrange"], c["yrange"], c["mesh"])
        assert np.allclose(x[::8], c["x"]) and np.allclose(Fx[::8], c["Fx"], rtol=1e-10, atol=1e-9), name
        assert np.allclose(y[::8], c["y"]) and np.allclose(Fy[::8], c["Fy"], rtol=1e-10, atol=1e-9), name


def test_fes_landmarks_double_well():
    # SURVEY Appendix A.5
    x, Fx = fes.projection_x(potentials.DOUBLE_WELL_2D, 300, [-2.25, 2.25], [-2, 2], 200)
    assert abs(x[np.argmin(Fx)] + 1.4585) < 0.02
    right = (x > 0.5)
    assert abs(Fx[right].min() - 7.934) < 0.01
    mid = (x > -0.5) & (x < 0.5)
    assert abs(Fx[mid].max() - 20.186) < 0.01
    assert abs(Fx[-1] - 47.94) < 0.01


def test_philox_known_answers():
    # Random123 kat_vectors
    o = philox.philox4x32_10(0, 0, 0, 0, 0, 0)
    assert [int(v) for v in o] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    f = 0xffffffff
    o = philox.philox4x32_10(f, f, f, f, f, f)
    assert [int(v) for v in o] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    o = philox.philox4x32_10(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0)
    assert [int(v) for v in o] == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_noise_statistics():
    z = np.concatenate([philox.dynamics_noise(7, np.arange(20000), t) for t in range(8)])
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1) < 0.01
    assert abs(np.mean(z[:, 0] * z[:, 1])) < 0.01
    assert abs(np.mean(z ** 4) - 3) < 0.1
    # consecutive steps of one walker are uncorrelated, and streams differ
    a = philox.dynamics_noise(7, np.arange(20000), 0)
    b = philox.dynamics_noise(7, np.arange(20000), 1)
    assert abs(np.mean(a[:, 0] * b[:, 0])) < 0.03
    assert not np.allclose(philox.velocity_noise(7, np.arange(4))[:, :2], a[:4])


def test_langevin_constants_and_equilibrium():
    k = langevin.constants()
    # SURVEY Appendix A.1
    assert abs(k["a"] - 0.8187307530779818) < 1e-15
    assert abs(k["b"] - 0.00906346234610091) < 1e-15
    assert abs(k["c"] - 0.9068260560644185) < 1e-13
    assert abs(k["sigma_v"] - 1.5793475822) < 1e-9
    # harmonic well U = x^2 + y^2 started in equilibrium: <x^2> = kT/2 and <v^2> = kT/m stay put
    n = 20000
    x = np.random.default_rng(0).standard_normal((n, 2)) * np.sqrt(k["kT"] / 2)
    v = langevin.maxwell_boltzmann(1, 0, n, k, 2)
    f = lambda p: potentials.force(potentials.SINGLE_WELL_2D, p)
    x, v = langevin.run_philox(x, v, f, k, seed=1, gid0=0, step0=0, nsteps=300)
    assert abs(np.mean(x ** 2) / (k["kT"] / 2) - 1) < 0.04
    assert abs(np.mean(v ** 2) / k["kT"] - 1) < 0.04


def test_langevin_chunking_is_invariant():
    k = langevin.constants()
    f = lambda p: potentials.force(potentials.DOUBLE_WELL_2D, p)
    x0 = np.array([[-1.4, 0.1], [1.3, -0.2], [0.0, 0.0]])
    v0 = np.zeros_like(x0)
    xa, va = langevin.run_philox(x0, v0, f, k, 5, 10, 0, 37)
    xb, vb = langevin.run_philox(x0, v0, f, k, 5, 10, 0, 13)
    xb, vb = langevin.run_philox(xb, vb, f, k, 5, 10, 13, 24)
    assert np.array_equal(xa, xb) and np.array_equal(va, vb)


def test_mc_placement_rule():
    beta = 1 / (langevin.KB_PY * 300)
    e = lambda x, y: potentials.energy_grad_xy(potentials.DOUBLE_WELL_2D, x, y)[0]
    pos, ok = langevin.mc_placement(3, 0, 5000, e, beta, (-2.5, 2.5, -2, 2))
    assert ok.all()
    assert np.all(e(pos[:, 0], pos[:, 1]) <= np.log(2) / beta + 1e-12)


def test_textbook_ves_fixed_point():
    """At V = -F the biased ensemble is uniform, so the textbook gradient vanishes (unpinned: literature)."""
    rng = np.random.default_rng(0)
    s = rng.uniform(-1, 1, 200000)
    P, _ = legendre.legendre_table(s, 6)
    sw, sf, sff = ves.moments(P.T)
    ft = np.zeros(7)
    ft[0] = 1.0
    grad, hess = ves.gradient_hessian(sw, sf, sff, ft, beta=0.4)
    assert np.all(np.abs(grad) < 0.01)
    # Cov of Legendre polynomials under the uniform density: diag 1/(2k+1), k >= 1
    assert np.allclose(np.diag(hess)[1:] / 0.4, 1 / (2 * np.arange(1, 7) + 1), atol=0.01)
    opt = ves.AveragedSGD(np.zeros(7), mu=0.5)
    ab = opt.step(grad, hess)
    assert ab.shape == (7,)


def test_mlp_param_grad_matches_finite_difference():
    rng = np.random.default_rng(0)
    sizes = [5, 4, 3]
    theta = rng.standard_normal(mlp.n_params(sizes))
    s = rng.uniform(0, 1, 7)
    g = mlp.param_grad(s, theta, sizes, row_weights=np.arange(7.0))
    h = 1e-6
    for i in range(len(theta)):
        tp = theta.copy(); tp[i] += h
        tm = theta.copy(); tm[i] -= h
        fd = np.dot(np.arange(7.0), (mlp.forward(s, tp, sizes)[0] - mlp.forward(s, tm, sizes)[0]) / (2 * h))
        assert abs(fd - g[i]) < 1e-5
