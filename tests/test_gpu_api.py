"""GPU tests of the reference-facing Python API (pyves_b200.basis / bias / langevin_dynamics / fes).

These read like the reference's own usage: the unit tests of tests/tests_unit/test_basis.py, the
fitting loop of examples/double_well_2D/static_legendre_double_well.py:53-73, the update() of
ves/bias.py:206-253 (golden vectors produced by the unmodified reference), and the run loop of
ves/langevin_dynamics.py:109-195.
"""
import contextlib
import io
import os
import pickle

import numpy as np
import pytest
from scipy.special import legendre as scipy_legendre

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import fes as ofes, langevin, legendre, mlp, potentials as opot  # noqa: E402
from pyves_b200 import basis, bias, fes, potentials, target  # noqa: E402
from pyves_b200.langevin_dynamics import SingleParticleSimulation  # noqa: E402

potentials._Potential.quiet = True


def quiet(fn, *a, **k):
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        out = fn(*a, **k)
    return out, buf.getvalue()


# ------------------------------------------------------------------ reference unit tests (test_basis.py)
def test_reference_unit_tests_legendre_polynomials():
    for device in ("cpu", "cuda"):
        x = torch.linspace(-1, 1, 1000, dtype=torch.float64, device=device)
        for degree in range(5):                                       # test_basis.py:11-25
            lves = basis.LegendreBasis1D.legendre_polynomial(x, degree).detach().cpu().numpy()
            assert np.allclose(lves, scipy_legendre(degree)(x.cpu().numpy()))
        weights = torch.rand(6, dtype=torch.float64)                  # test_basis.py:28-45
        bench = sum(weights[i] * basis.LegendreBasis1D.legendre_polynomial(x, i).cpu() for i in range(6))
        out = basis.LegendreBasis1D.legendre_polynomial_expansion(x, 5, weights).detach().cpu()
        assert np.allclose(bench.numpy(), out.numpy())
        # the path-CV class carries the same two helpers upstream (ves/basis.py:442-507)
        out2 = basis.LegendreBasis2DPathCV.legendre_polynomial_expansion(x, 5, weights).detach().cpu()
        assert np.array_equal(out2.numpy(), out.numpy())
        assert np.allclose(basis.LegendreBasis2DPathCV.legendre_polynomial(x, 3).cpu().numpy(), scipy_legendre(3)(x.cpu().numpy()))
        for degree in range(1, 5):                                   # test_basis.py:48-68
            xg = torch.linspace(-1, 1, 1000, dtype=torch.float64, device=device, requires_grad=True)
            l = basis.LegendreBasis1D.legendre_polynomial(xg, degree)
            g = torch.autograd.grad(l, xg, grad_outputs=torch.ones_like(l))[0].cpu().numpy()
            assert np.allclose(g, scipy_legendre(degree).deriv()(xg.detach().cpu().numpy()))
        # the assertion test_basis.py:70-89 meant to make: gradient of the expansion w.r.t. x
        xg = torch.linspace(-1, 1, 1000, dtype=torch.float64, device=device, requires_grad=True)
        o = basis.LegendreBasis1D.legendre_polynomial_expansion(xg, 5, weights)
        g = torch.autograd.grad(o, xg, grad_outputs=torch.ones_like(o))[0].cpu().numpy()
        ref = sum(float(weights[i]) * scipy_legendre(i).deriv()(xg.detach().cpu().numpy()) for i in range(6))
        assert np.allclose(g, ref)


def test_modules_forward_backward_vs_golden(golden):
    for c in golden("legendre1d.json"):
        m = basis.LegendreBasis1D(c["degree"], c["min"], c["max"], axis=c["axis"], weights=np.array(c["weights"]))
        for device in ("cuda", "cpu"):
            pos = torch.tensor(c["positions"], dtype=torch.float64, device=device, requires_grad=True)
            E = m(pos)
            assert E.shape == (len(c["E"]),) and E.dtype == torch.float64 and E.device.type == device
            E.sum().backward()
            assert np.allclose(E.detach().cpu().numpy(), c["E"], rtol=1e-9, atol=1e-9), c["tag"]
            assert np.allclose(-pos.grad.cpu().numpy(), c["F"], rtol=1e-9, atol=1e-8), c["tag"]
            # d(sum E)/dw_k = sum_rows P_k(s)
            ref = legendre.basis1d(np.array(c["positions"]), c["degree"], c["min"], c["max"], c["axis"]).sum(0)
            assert np.allclose(m.weights.grad.numpy(), ref, rtol=1e-9, atol=1e-9)
            m.weights.grad = None
        # float32 positions give float32 energies
        assert m(torch.tensor(c["positions"], dtype=torch.float32, device="cuda")).dtype == torch.float32
    for c in golden("nnbasis1d.json"):
        act = torch.nn.ReLU if c["act"] == "relu" else torch.nn.Tanh
        m = basis.NNBasis1D(c["min"], c["max"], axis=c["axis"], hidden_layer_sizes=c["sizes"], act=act)
        o = 0
        with torch.no_grad():
            for p in m.parameters():
                p.copy_(torch.tensor(c["theta"][o:o + p.numel()], dtype=torch.float64).reshape(p.shape))
                o += p.numel()
        pos = torch.tensor(c["positions"], dtype=torch.float64, device="cuda", requires_grad=True)
        E = m(pos)
        gw = torch.arange(1.0, len(c["E"]) + 1, dtype=torch.float64, device="cuda")
        (E * gw).sum().backward()
        assert np.allclose(E.detach().cpu().numpy(), c["E"], rtol=1e-9, atol=1e-9), c["tag"]
        assert np.allclose(-pos.grad.cpu().numpy() / gw.cpu().numpy()[:, None], c["F"], rtol=1e-9, atol=1e-9), c["tag"]
        a = {'x': 0, 'y': 1}[c["axis"]]
        s = (np.array(c["positions"])[:, a] - c["min"]) / (c["max"] - c["min"])
        ref = mlp.param_grad(s, c["theta"], c["sizes"], {"relu": 0, "tanh": 1}[c["act"]], row_weights=gw.cpu().numpy())
        got = np.concatenate([p.grad.numpy().ravel() for p in m.parameters()])
        assert np.allclose(got, ref, rtol=1e-8, atol=1e-8), c["tag"]
        # the reference's fitting shape [M, 1, 1] -> [M, 1] (deepves example :49-62)
        x3 = torch.tensor(np.array(c["positions"])[:, a], dtype=torch.float64).unsqueeze(-1).unsqueeze(-1)
        if c["axis"] == 'x':
            assert m(x3).shape == (len(c["E"]), 1)
    for c in golden("pathcv.json"):
        m = basis.LegendreBasis2DPathCV(c["degree"], np.array(c["x_i"]), np.array(c["y_i"]), c["lam"], np.array(c["weights"]), c["phi"])
        pos = torch.tensor(c["positions"], dtype=torch.float64, device="cuda", requires_grad=True)
        E = m(pos)
        E.sum().backward()
        assert np.allclose(E.detach().cpu().numpy(), c["E"], rtol=1e-8) and np.allclose(-pos.grad.cpu().numpy(), c["F"], rtol=1e-7, atol=1e-7)
    c = golden("harmonic.json")
    hm = bias.HarmonicBias_SingleParticle_1D_ForceModule(c["k"], c["s0"], axis=c["axis"])
    pos = torch.tensor(c["positions"], dtype=torch.float64, device="cuda", requires_grad=True)
    E = hm(pos)
    E.backward()
    assert np.allclose(float(E), c["E"][0]) and np.allclose(-pos.grad.cpu().numpy(), c["F"])


def test_fit_legendre_to_free_energy_like_the_example():
    """static_legendre_double_well.py:32-73: Adam fit of LegendreBasis1D to the analytic F(x) on [M, 1] inputs."""
    pot = potentials.DoubleWellPotential2D()
    x, Fx = fes.projection_x(pot, 300, [-2.25, 2.25], [-2, 2])
    xr, Fr = ofes.projection_x(opot.DOUBLE_WELL_2D, 300, [-2.25, 2.25], [-2, 2])
    assert np.array_equal(x, xr) and np.allclose(Fx, Fr, rtol=1e-12)
    kT = 8.3145 / 1000 * 300
    xt = torch.tensor(x, device="cuda").unsqueeze(-1)
    Ft = torch.tensor(Fx * kT, device="cuda")
    fn = basis.LegendreBasis1D(5, -2.5, 2.5, weights=np.zeros(6))
    opt = torch.optim.Adam(fn.parameters(), lr=0.5)
    loss_fn = torch.nn.MSELoss()
    first = None
    for step in range(400):
        opt.zero_grad()
        loss = loss_fn(fn(xt), Ft)
        first = first if first is not None else float(loss)
        loss.backward()
        opt.step()
    assert float(loss) < 1e-3 * first
    # deterministic limit of that fit: least squares (F_x is exactly quartic)
    ref = np.polynomial.legendre.legfit(x / 2.5, Fx * kT, 5)
    assert np.allclose(fn.weights.detach().numpy(), ref, atol=0.5)


def test_reference_update_known_answers(golden, tmp_path):
    """VESBias_SingleParticle_1D.update against the unmodified reference: printed loss, gradients, new weights, files."""
    for c in golden("update.json"):
        if c["tag"].startswith("nn"):
            torch.manual_seed(3)
            model = basis.NNBasis1D(-4.0, 4.0, hidden_layer_sizes=[8, 8])
        else:
            model = basis.LegendreBasis1D(5, -2.5, 2.5, weights=np.array(c["theta0"]))
        theta0 = np.concatenate([p.detach().numpy().ravel() for p in model.parameters()])
        assert np.allclose(theta0, c["theta0"], rtol=0, atol=1e-15), c["tag"]        # same torch init stream as upstream
        d = tmp_path / c["tag"]
        d.mkdir()
        b = bias.VESBias_SingleParticle_1D(model, target.Target_Uniform_HardSwitch_x(c["mesh"]), c["beta"], c["optimizer"],
                                           c["optimizer_params"], str(d / "model.pt"), update_steps=c["update_steps"])
        traj = np.array(c["traj"])
        for u in range(c["nupdates"]):
            tr = traj[: len(traj) - (c["nupdates"] - 1 - u) * 20]
            _, out = quiet(b.update, tr)
            lines = out.strip().splitlines()
            assert lines[0] == "{} {}".format(max(len(tr) - c["update_steps"], 0), len(tr))      # ves/bias.py:218
            assert abs(float(lines[1].split("=")[1]) - c["loss_printed"][u]) < 2e-6, c["tag"]
            grad = np.concatenate([p.grad.numpy().ravel() for p in model.parameters()])
            assert np.allclose(grad, c["grad"][u], rtol=1e-7, atol=1e-9), c["tag"]
            theta = np.concatenate([p.detach().numpy().ravel() for p in model.parameters()])
            assert np.allclose(theta, c["theta"][u], rtol=1e-7, atol=1e-9), c["tag"]
        assert sorted(os.listdir(d)) == c["files"], c["tag"]
        reloaded = torch.jit.load(str(d / "model.pt"))
        pos = torch.tensor([[0.3, 0.0, 0.0]], dtype=torch.float64)
        assert np.allclose(reloaded(pos).detach().numpy(), model(pos).detach().numpy(), rtol=1e-9)


# ------------------------------------------------------------------ the simulation driver
def test_single_walker_run_reference_format(tmp_path):
    """n_walkers = 1: reference file formats, frame/energy cadence, checkpoint + resume, and the trajectory itself
    against the oracle driven by the same Philox stream (dims = 3, FP64)."""
    pot = potentials.DoubleWellPotential2D()
    w = np.array([1.0, -2.0, 0.5, 0.1, -0.3, 0.2])
    m = basis.LegendreBasis1D(5, -2.5, 2.5, weights=w)
    init = np.array([[-1.3, 0.2, 0.0]])
    sim, out = quiet(SingleParticleSimulation, pot, init_coord=init, seed=42)
    b = bias.StaticBias_SingleParticle(m, str(tmp_path / "bias.pt"))
    sim.init_ves(b, static=True)
    v0 = sim.get_state().velocities.copy()
    _, log = quiet(sim, nsteps=100, chkevery=40, trajevery=1, energyevery=1, chkfile=str(tmp_path / "chk.pkl"),
                   trajfile=str(tmp_path / "traj.dat"), energyfile=str(tmp_path / "energies.dat"))
    assert "[VES] StaticBias_SingleParticle bias: Initializing static bias." in log
    assert log.count("Checkpoint at") == 3 and "Checkpoint at  0.9900000 ps" in log
    assert sim.traj.shape == (100, 3) and np.allclose(sim.traj[0], init[0])
    assert len(sim.PE) == 100 and len(sim.KE) == 100
    lines = open(tmp_path / "traj.dat").read().splitlines()
    assert lines[0] == "# t[ps]    x [nm]    y [nm]    z[nm]" and len(lines) == 101
    assert lines[1] == "{:10.5f}\t{:10.7f}\t{:10.7f}\t{:10.7f}".format(0.0, -1.3, 0.2, 0.0)
    assert lines[100].startswith("{:10.5f}".format(0.99))
    el = open(tmp_path / "energies.dat").read().splitlines()
    assert el[0] == "# t[ps]    PE [kJ/mol]    KE [kJ/mol]" and len(el) == 101
    # oracle: same seed, walker id 0, 3 simulated dimensions
    k = langevin.constants()
    f = lambda x: opot.force(opot.DOUBLE_WELL_2D, x) + legendre.legendre1d(x, 5, -2.5, 2.5, 'x', w)[1]
    assert np.allclose(v0, langevin.maxwell_boltzmann(42, 0, 1, k, 3), rtol=1e-12)
    xr, vr, frames = None, None, []
    x, v = init.copy(), v0.copy()
    from oracle import philox
    for t in range(100):
        frames.append(x[0].copy())
        x, v = langevin.step(x, v, f, k, philox.dynamics_noise(42, [0], t, dims=3))
    assert np.allclose(sim.traj, np.array(frames), rtol=1e-9, atol=1e-9)
    U0 = opot.energy(opot.DOUBLE_WELL_2D, init) + legendre.legendre1d(init, 5, -2.5, 2.5, 'x', w)[0]
    assert abs(sim.PE[0] - U0[0]) < 1e-9 and abs(sim.KE[0] - langevin.kinetic_energy(v0, f(init), k)[0]) < 1e-9
    # resume from the last checkpoint: continuing gives the same trajectory as an uninterrupted run
    st = pickle.load(open(tmp_path / "chk.pkl", "rb"))
    assert st.step == 100 and st.getPositions().shape == (1, 3)
    sim2, _ = quiet(SingleParticleSimulation, pot, init_state=st, seed=42)
    sim2.init_ves(b, static=True)
    quiet(sim2, nsteps=20, chkevery=0, trajfile=None, energyfile=None, chkfile=None)
    for t in range(100, 120):
        frames.append(x[0].copy())
        x, v = langevin.step(x, v, f, k, philox.dynamics_noise(42, [0], t, dims=3))
    assert np.allclose(sim2.traj, np.array(frames[100:]), rtol=1e-9, atol=1e-9)


def test_dynamic_bias_update_cadence(tmp_path):
    """First update at i == startafter, then whenever i % learnevery == 0 (ves/langevin_dynamics.py:134-159);
    update(traj) sees i frames."""
    pot = potentials.SzaboBerezhkovskiiPotential()
    model = basis.LegendreBasis1D(4, -4.0, 4.0, weights=np.zeros(5))
    seen = []

    class Spy(bias.VESBias_SingleParticle_1D):
        def update(self, traj):
            seen.append(traj.shape[0])
            super().update(traj)

    b = Spy(model, target.Target_Uniform_HardSwitch_x(21), 1 / (8.3145e-3 * 300), "Adam", {"lr": 0.01},
            str(tmp_path / "m.pt"), update_steps=30)
    sim, _ = quiet(SingleParticleSimulation, pot, init_coord=np.array([[2.2, 2.2, 0.0]]), seed=1)
    sim.init_ves(b, startafter=20, learnevery=15)
    _, log = quiet(sim, nsteps=61, chkevery=1000, chkfile=str(tmp_path / "c.pkl"), trajfile=str(tmp_path / "t.dat"),
                   energyfile=None)
    assert seen == [20, 30, 45, 60]
    assert "Initializing dynamic bias at timestep 20." in log and "Updating dynamic bias at timestep 45." in log
    assert sorted(os.listdir(tmp_path)) == ["c.pkl", "m.pt", "m.pt.iter20", "m.pt.iter30", "m.pt.iter45", "m.pt.iter60", "t.dat"]
    assert sim.traj.shape == (61, 3)
    assert not np.allclose(model.weights.detach().numpy(), 0)


@pytest.mark.parametrize("device_update", [False, True], ids=["host_update", "device_update"])
def test_ensemble_ves_converges_to_analytic_fes(device_update):
    """Textbook VES with an ensemble: LegendreBasis1D on the 2D double well, uniform target, averaged SGD.
    Converged F(x) = -V(x) agrees with the analytic projection (ves/visualization.py:185-201) within 0.1 kT;
    with the host-side optimiser and with the device-resident update loop of the driver."""
    pot = potentials.DoubleWellPotential2D()
    kT = 8.3145e-3 * 300
    model = basis.LegendreBasis1D(6, -2.5, 2.5, weights=np.zeros(7))
    b = bias.VESBias_Ensemble(model, target.Target_Uniform_HardSwitch_x(401), 1 / kT, "averaged_sgd",
                              {"mu": 3.0, "second_order": True, "averaging": "ewma", "decay": 0.9})
    sim, _ = quiet(SingleParticleSimulation, pot, n_walkers=65536, seed=7, temp=300, device_update=device_update)
    sim.place_walkers(xmin=-2.5, xmax=2.5, ymin=-2, ymax=2)
    sim.init_ves(b, startafter=100, learnevery=100)
    quiet(sim, nsteps=60001, chkevery=0, trajevery=0, energyevery=0, chkfile=None, trajfile=None, energyfile=None)
    assert b.n_updates == 600
    x = np.linspace(-2.25, 2.25, 200)
    V = model(torch.tensor(np.stack([x, 0 * x], 1), device="cuda")).detach().cpu().numpy()
    F_est = (-V - (-V).min()) / kT
    _, F_ref = fes.projection_x(pot, 300, [-2.25, 2.25], [-2, 2])
    rmse = fes.rmse_kt(F_est, F_ref, clip=20.0)
    assert rmse < 0.1, rmse
    # the biased ensemble is flat along x: histogram FES of the walkers varies by well under 1 kT
    centres, bF = fes.histogram_fes_1d(sim.ens, (-2.0, 2.0), 40, axis=0)
    assert bF.max() < 0.5


def _device_vs_host(make_bias, pot_kind, pot_params, box, cov, n_updates=24, stride=20, precision="fp32"):
    from pyves_b200 import _lib
    from pyves_b200.bias import DeviceUpdater
    kT = 8.3145e-3 * 300
    runs = []
    for mode in ("host", "device"):
        b = make_bias()
        ens = _lib.Ensemble(30000, dims=2, seed=7, precision=precision)
        ens.set_integrator(1.0, 8.31446261815324e-3 * 300, 20.0, 0.01)
        ens.set_potential(pot_kind, pot_params)
        assert ens.init_positions_mc(box, 1 / kT, ntrials=2000) == 0
        ens.init_velocities()
        if mode == "host":
            b.force.apply(ens)
            for _ in range(n_updates):
                ens.step(stride)
                sw, sf, sff = ens.moments(cov=cov)
                b.update_from_moments(sw, sf, sff)
                b.force.apply(ens)
        else:
            upd = DeviceUpdater(b, ens)
            for _ in range(n_updates):
                ens.step(stride)
                sw, sf, sff = ens.moments(cov=cov)
                upd.update(sw, sf, sff)
            E_dev, _ = ens.eval_bias(ens.get_state()[0].T.contiguous()[:64], want_force=False)   # by-value consumer in device mode
            upd.sync_to_host()
            upd.close()
            b.force.apply(ens)                                                                   # back to host mode
            E_host, _ = ens.eval_bias(ens.get_state()[0].T.contiguous()[:64], want_force=False)
            assert torch.allclose(E_dev, E_host, rtol=1e-5, atol=1e-5)      # same coefficients either way
        x, _ = ens.get_state(host=True)
        runs.append((b.optimizer.abar.copy(), b.optimizer.alpha.copy(), np.asarray(b.target.p).copy(), x, b.n_updates))
        ens.close()
    (a0, al0, p0, x0, n0), (a1, al1, p1, x1, n1) = runs
    assert n0 == n1 == n_updates
    scale = np.max(np.abs(a0))
    assert scale > 0.05
    assert np.max(np.abs(a0 - a1)) < 1e-9 * scale and np.max(np.abs(al0 - al1)) < 1e-9 * np.max(np.abs(al0))
    assert np.max(np.abs(p0 - p1)) < 1e-9 * np.max(p0)
    if precision == "fp32":
        assert np.array_equal(x0, x1)        # the 1e-16 differences of the FP64 coefficients vanish in the FP32 cast
    else:
        assert np.max(np.abs(x0 - x1)) < 1e-9


def test_device_resident_update_matches_host_update():
    """DeviceUpdater (optimiser, well-tempered target refresh and coefficient hand-over all on the device,
    pyves_set_bias_legendre{1d,2d}_device) reproduces VESBias_Ensemble.update_from_moments: same coefficients,
    target and walkers (bit for bit) after 24 updates -- static 16 x 16 kernels (weights in the __constant__
    array), a runtime-size 2D basis (device buffer), a 1D basis with the second-order step (configs[2] shape)
    and a 1D well-tempered target."""
    from pyves_b200 import _lib
    kT = 8.3145e-3 * 300
    sgd = {"mu": 0.5, "averaging": "ewma", "decay": 0.8, "precondition": 1.0}
    for deg, prec in ((15, "fp32"), (9, "fp32"), (15, "fp64"), (7, "fp64")):
        _device_vs_host(lambda: bias.VESBias_Ensemble(basis.LegendreBasis2D(deg, deg, -2.5, 2.5, -2.0, 2.0),
                                                      target.Target_WellTempered_2D(40, 40, 10.0), 1 / kT, "averaged_sgd", sgd,
                                                      target_update_every=5),
                        _lib.POT_DOUBLE_WELL_2D, (), (-2.5, 2.5, -2.0, 2.0), False, precision=prec)
    _device_vs_host(lambda: bias.VESBias_Ensemble(basis.LegendreBasis1D(12, -3, 6, 'x', weights=np.zeros(13)), target.Target_Uniform_HardSwitch_x(200), 1 / kT,
                                                  "averaged_sgd", {"mu": 0.5, "second_order": True, "averaging": "ewma", "decay": 0.8},
                                                  target_update_every=0),
                    _lib.POT_SLIP_BOND_2D, (0, 0, 1, 5, 4, 0, 2), (-8, 10, -6, 8), True)
    _device_vs_host(lambda: bias.VESBias_Ensemble(basis.LegendreBasis1D(12, -3, 6, 'x', weights=np.zeros(13)), target.Target_Uniform_HardSwitch_x(200), 1 / kT,
                                                  "averaged_sgd", {"mu": 0.5, "second_order": True, "averaging": "ewma", "decay": 0.8},
                                                  target_update_every=0),
                    _lib.POT_SLIP_BOND_2D, (0, 0, 1, 5, 4, 0, 2), (-8, 10, -6, 8), True, precision="fp64")
    _device_vs_host(lambda: bias.VESBias_Ensemble(basis.LegendreBasis1D(5, -2.5, 2.5, 'x', weights=np.zeros(6)), target.Target_WellTempered_x(100, 5.0), 1 / kT,
                                                  "averaged_sgd", {"mu": 0.5, "averaging": "running"}, target_update_every=4),
                    _lib.POT_DOUBLE_WELL_2D, (), (-2.5, 2.5, -2.0, 2.0), False)


def test_chunked_handles_async_copies_match_single_handle():
    """The e2e pipeline of bench.py: the ensemble cut into handles on their own streams, state copied from / to
    pinned host memory asynchronously (on_host = 2), one DeviceUpdater feeding all handles.  Walkers and
    coefficients equal the single-handle run to rounding."""
    from pyves_b200 import _lib
    from pyves_b200.bias import DeviceUpdater
    kT = 8.3145e-3 * 300
    n, nch, stride, iters = 40000, 4, 30, 6

    def make_bias():
        return bias.VESBias_Ensemble(basis.LegendreBasis2D(15, 15, -2.5, 2.5, -2.0, 2.0), target.Target_WellTempered_2D(40, 40, 10.0),
                                     1 / kT, "averaged_sgd", {"mu": 0.5, "averaging": "ewma", "decay": 0.8, "precondition": 1.0},
                                     target_update_every=3)

    def make_ens(count, offset):
        e = _lib.Ensemble(count, dims=2, seed=11, precision="fp32")
        e.set_integrator(1.0, 8.31446261815324e-3 * 300, 20.0, 0.01)
        e.set_potential(_lib.POT_DOUBLE_WELL_2D)
        e.set_walker_offset(offset)
        return e

    ref = make_ens(n, 0)
    assert ref.init_positions_mc((-2.5, 2.5, -2.0, 2.0), 1 / kT) == 0
    ref.init_velocities()
    x0, v0 = ref.get_state(host=True)
    b_ref = make_bias()
    u_ref = DeviceUpdater(b_ref, ref)
    for _ in range(iters):
        ref.step(stride)
        sw, sf, _ = ref.moments()
        u_ref.update(sw, sf)
    xr, vr = ref.get_state(host=True)
    u_ref.sync_to_host()

    cn = n // nch
    subs = [make_ens(cn, c * cn) for c in range(nch)]
    streams = [torch.cuda.Stream() for _ in range(nch)]
    hp = [torch.from_numpy(np.ascontiguousarray(x0[:, c * cn:(c + 1) * cn])).pin_memory().numpy() for c in range(nch)]
    hv = [torch.from_numpy(np.ascontiguousarray(v0[:, c * cn:(c + 1) * cn])).pin_memory().numpy() for c in range(nch)]
    b = make_bias()
    upd = DeviceUpdater(b, subs[0])
    for e in subs[1:]:
        upd.add_handle(e)
    main = torch.cuda.current_stream()
    # per-chunk moment buffers allocated once on the main stream and written through out=: nothing allocated on a side
    # stream is consumed on the main stream (the caching allocator could recycle such a block too early)
    mom = [(torch.zeros(1, dtype=torch.float64, device="cuda"), torch.zeros(256, dtype=torch.float64, device="cuda"), None) for _ in range(nch)]
    for _ in range(iters):
        for c in range(nch):
            streams[c].wait_stream(main)
            with torch.cuda.stream(streams[c]):
                if c % 2:                                   # the same four operations as one C call (pyves_iterate_pinned)
                    subs[c].iterate_pinned(hp[c], hv[c], stride, mom[c])
                else:
                    subs[c].set_state(hp[c], hv[c], async_copy=True)
                    subs[c].step(stride)
                    subs[c].moments(out=mom[c])
                    subs[c].get_state(out=(hp[c], hv[c]), async_copy=True)
        for c in range(nch):
            main.wait_stream(streams[c])
        upd.update(sum(m[0] for m in mom), sum(m[1] for m in mom))
        torch.cuda.synchronize()
    upd.sync_to_host()
    # the chunk moments are added in another order than the single-handle reduction (FP64, ~1e-16 relative), so a
    # coefficient can round to the neighbouring FP32 value: walkers agree to FP32 rounding, not bit for bit
    assert np.max(np.abs(np.concatenate(hp, axis=1) - xr)) < 1e-4 and np.max(np.abs(np.concatenate(hv, axis=1) - vr)) < 1e-3
    assert np.max(np.abs(b.optimizer.abar - b_ref.optimizer.abar)) < 1e-7 * np.max(np.abs(b_ref.optimizer.abar))
    for h in subs + [ref]:
        h.close()


def test_step_wave_walkers():
    """pyves_step_wave_walkers: SMs x walkers per CTA for the tensor-core step kernels (one CTA per SM), 0 for the kernels with
    many small CTAs and for FP64 handles."""
    from pyves_b200 import _lib
    n_sm = torch.cuda.get_device_properties(0).multi_processor_count

    def wave(prec, setup, variant=-1):
        e = _lib.Ensemble(1000, dims=2, seed=1, precision=prec)
        e.set_integrator(1.0, 2.494, 20.0, 0.01)
        e.set_potential(_lib.POT_DOUBLE_WELL_2D)
        setup(e)
        e.set_tuning(variant)
        w = e.wave_walkers()
        e.close()
        return w

    leg = lambda k: (lambda e: e.set_bias_legendre2d(k - 1, k - 1, (-2.5, 2.5, -2, 2), np.zeros(k * k)))
    assert wave("fp32", leg(16)) == n_sm * 512 and wave("fp32", leg(32)) == n_sm * 768 and wave("fp32", leg(64)) == n_sm * 512
    assert wave("fp32", leg(8)) == 0 and wave("fp32", leg(8), 7) == n_sm * 512      # static 8 x 8 double-well kernel unless forced
    assert wave("fp32", leg(16), 6) == 0 and wave("fp64", leg(16)) == 0
    assert wave("fp32", lambda e: e.set_bias_legendre1d(0, 5, -2.5, 2.5, np.zeros(6))) == 0
    torch.manual_seed(0)
    m = basis.NNBasis1D(min=-4, max=4, axis="x", hidden_layer_sizes=[48, 48, 48])
    theta = np.concatenate([p.detach().numpy().ravel() for p in m.parameters()])
    nn = lambda e: e.set_bias_mlp(0, -4.0, 4.0, [48, 48, 48], 0, theta)
    assert wave("fp32", nn) == 0 and wave("fp32", nn, 7) == n_sm * 640           # default: the piecewise-constant table, many small CTAs


def test_ensemble_ves_along_path_cv():
    """Ensemble VES with LegendreBasis2DPathCV (ves/basis.py:386-548) and a uniform target in the scaled path
    coordinate: the path runs along the x axis of the 2D double well through both minima, so the converged bias
    must flatten the 20 kT barrier -- the walkers, all started in one well, end up spread over the whole path, and
    F(s) = -V(s) reproduces the analytic x projection (s is linear in x for this path)."""
    pot = potentials.DoubleWellPotential2D()
    kT = 8.3145e-3 * 300
    n_img = 40
    x_i, y_i = np.linspace(-2.2, 2.2, n_img), np.zeros(n_img)
    model = basis.LegendreBasis2DPathCV(8, x_i, y_i, 100.0, weights=np.zeros(9))
    b = bias.VESBias_Ensemble(model, target.Target_Uniform_HardSwitch_x(401), 1 / kT, "averaged_sgd",
                              {"mu": 3.0, "second_order": True, "averaging": "ewma", "decay": 0.9})
    init = np.array([[-1.47, 0.0, 0.0]])
    sim, _ = quiet(SingleParticleSimulation, pot, n_walkers=32768, seed=5, temp=300, init_coord=init)
    sim.init_ves(b, startafter=100, learnevery=100)
    quiet(sim, nsteps=60001, chkevery=0, trajevery=0, energyevery=0, chkfile=None, trajfile=None, energyfile=None)
    assert b.n_updates == 600
    pos = sim.get_state().positions
    assert (pos[:, 0] > 0.5).mean() > 0.25 and (pos[:, 0] < -0.5).mean() > 0.25      # crossed the barrier, both wells populated
    # F(s) against the analytic projection on x(s): s = (i + 1) / N at image i  ->  x = x_0 + (N s - 1) dx
    F = b.fes() / kT
    sp = np.asarray(b.target.x)
    xs = x_i[0] + (n_img * (sp + 1) / 2 - 1) * (x_i[1] - x_i[0])
    xa, F_ref = fes.projection_x(pot, 300, [-2.25, 2.25], [-2, 2])
    sel = (xs > -1.9) & (xs < 1.9)
    F_ref_s = np.interp(xs[sel], xa, F_ref)
    d = (F[sel] - F[sel].min()) - (F_ref_s - F_ref_s.min())
    keep = F_ref_s - F_ref_s.min() < 15
    assert np.sqrt(np.mean(d[keep] ** 2)) < 1.0, np.sqrt(np.mean(d[keep] ** 2))


def test_example_script_ensemble_ves():
    """examples/double_well_2D_ensemble_ves.py (configs[1] through the `ves.*` names, device-resident updates):
    converged 2D surface within 0.1 kT of the analytic one at test size."""
    import importlib.util
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples", "double_well_2D_ensemble_ves.py")
    spec = importlib.util.spec_from_file_location("example_ensemble_ves", path)
    mod = importlib.util.module_from_spec(spec)
    with contextlib.redirect_stdout(io.StringIO()):
        spec.loader.exec_module(mod)
        rmse = mod.main(200000, 60000, quiet=True)
    assert rmse < 0.1, rmse
