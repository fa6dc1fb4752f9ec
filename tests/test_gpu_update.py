"""GPU tests of the device-resident update loop beyond the Legendre / averaged-SGD case of test_gpu_api.py:
NN parameters handed over on the device (pyves_set_bias_mlp_device), the torch.optim rules as one kernel
(pyves_optimizer_step; reference semantics ves/bias.py:193-200,236-245), the empty-average guard, stream ordering of the
host-fed parameter uploads, and the checkpoint of the whole update state (reference hooks
ves/langevin_dynamics.py:75-76,197-204, ves/bias.py:249-253)."""
import contextlib
import io
import pickle

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from pyves_b200 import _lib, basis, bias, potentials, target  # noqa: E402
from pyves_b200.bias import DeviceUpdater  # noqa: E402
from pyves_b200.langevin_dynamics import SingleParticleSimulation  # noqa: E402

potentials._Potential.quiet = True
KT = 8.3145e-3 * 300
KB = 8.31446261815324e-3


def quiet(fn, *a, **k):
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        return fn(*a, **k)


def sb_ensemble(n, precision, seed=7):
    e = _lib.Ensemble(n, dims=2, seed=seed, precision=precision)
    e.set_integrator(1.0, KB * 300, 20.0, 0.01)
    e.set_potential(_lib.POT_SZABO_BEREZHKOVSKII, (2.2, 4.0, 4.04, 4.84))
    assert e.init_positions_mc((-4, 4, -4, 4), 1 / KT, ntrials=2000) == 0
    e.init_velocities()
    return e


@pytest.mark.parametrize("sizes,act", [([8, 8], "relu"), ([48, 48, 48], "relu"), ([48, 48, 48], "tanh"), ([24, 16, 32], "tanh"),
                                       ([128, 128], "relu"), ([20, 20, 20, 20], "relu"), ([12, 96, 40], "relu")],
                         ids=lambda v: "x".join(map(str, v)) if isinstance(v, list) else v)
@pytest.mark.parametrize("precision", ["fp32", "fp64"])
def test_mlp_device_parameters_match_host_upload(sizes, act, precision):
    """pyves_set_bias_mlp_device (convert + fold + repack on the device) gives the step, eval and moment kernels exactly
    the parameters pyves_set_bias_mlp builds on the host: energies, forces, trajectories and gradient moments are equal
    bit for bit (every MLP functor: streamed two-layer, H = 32 / 48 / 64 matmul, generic)."""
    torch.manual_seed(5)
    m = basis.NNBasis1D(min=-4, max=4, axis='x', hidden_layer_sizes=sizes, act=torch.nn.ReLU if act == "relu" else torch.nn.Tanh)
    theta = np.concatenate([p.detach().numpy().ravel() for p in m.parameters()])
    th_dev = torch.tensor(theta, dtype=torch.float64, device="cuda")
    out = []
    for mode in ("host", "device"):
        e = sb_ensemble(3000, precision)
        if mode == "host":
            e.set_bias_mlp(0, -4.0, 4.0, sizes, m._act, theta)
        else:
            e.set_bias_mlp_device(0, -4.0, 4.0, sizes, m._act, th_dev)
        assert e.basis_size == theta.size
        pos = torch.tensor(np.random.default_rng(0).uniform(-5, 5, (500, 2)), dtype=e.torch_dtype, device="cuda")
        E, F = e.eval_bias(pos)
        e.step(40)
        x, v = e.get_state()
        sw, sf, _ = e.moments()
        out.append((E.clone(), F.clone(), x.clone(), v.clone(), sf.clone()))
        e.close()
    for a, b in zip(*out):
        assert torch.equal(a, b)
    with pytest.raises(NotImplementedError):
        e = sb_ensemble(10, "fp32")
        e.set_bias_mlp_device(0, -4.0, 4.0, [8], _lib.ACT_RELU, torch.zeros(25, dtype=torch.float64, device="cuda"))


OPT_CASES = [("Adam", {"lr": 1e-3}, "running"), ("Adam", {"lr": 2e-3, "betas": (0.8, 0.99), "eps": 1e-6}, None), ("Adam", {"lr": 1e-3}, 0.9),
             ("SGD", {"lr": 5e-3}, None), ("SGD", {"lr": 5e-3, "momentum": 0.7}, "running"), ("RMSprop", {"lr": 1e-3}, None),
             ("RMSprop", {"lr": 1e-3, "alpha": 0.9, "eps": 1e-7}, 0.8)]


@pytest.mark.parametrize("opt,params,gavg", OPT_CASES, ids=["%s-%s-%s" % (o, "-".join(sorted(p)), g) for o, p, g in OPT_CASES])
def test_device_resident_nn_update_matches_torch_optim(opt, params, gavg):
    """DeviceUpdater for NNBasis1D (pyves_optimizer_step + pyves_set_bias_mlp_device + target averages on the device)
    reproduces VESBias_Ensemble.update_from_moments, which steps the reference's torch.optim object on the host
    (ves/bias.py:193-200,236-245): parameters, optimiser state and gradient average after 4 updates agree to 1e-6
    (gradient averages to 1e-9), walkers to FP32 rounding.  (Not more updates: Adam and RMSprop divide by sqrt(v) + eps, so a parameter whose gradient
    sits at the eps level -- a ReLU unit that is almost dead -- turns a 1e-17 rounding difference into a tenfold larger
    one per update; measured with tools/diag/nn_update_diag.py: 6e-17 after 10 updates, 1e-5 after 20.  The kernel itself
    is pinned to torch.optim step by step in test_optimizer_step_kernel_matches_torch_optim.)"""
    n_upd = 4
    runs = []
    for mode in ("host", "device"):
        torch.manual_seed(2)
        model = basis.NNBasis1D(min=-4, max=4, axis='x', hidden_layer_sizes=[48, 48, 48])
        b = bias.VESBias_Ensemble(model, target.Target_Uniform_HardSwitch_x(200), 1 / KT, opt, dict(params), target_update_every=0,
                                  gradient_average=gavg)
        e = sb_ensemble(20000, "fp32")
        if mode == "host":
            b.force.apply(e)
            for _ in range(n_upd):
                e.step(20)
                sw, sf, _ = e.moments()
                b.update_from_moments(sw, sf)
                b.force.apply(e)
        else:
            assert DeviceUpdater.supports(b)
            upd = DeviceUpdater(b, e)
            for _ in range(n_upd):
                e.step(20)
                bufs = upd.moment_buffers()
                e.moments(out=bufs)
                upd.update(*bufs)
            upd.sync_to_host()
            upd.close()
        x, _ = e.get_state(host=True)
        theta = np.concatenate([p.detach().numpy().ravel() for p in model.parameters()])
        st = b.optimizer.state
        key = {"Adam": "exp_avg", "RMSprop": "square_avg", "SGD": "momentum_buffer"}[opt]
        s1 = np.concatenate([st[p][key].numpy().ravel() for p in model.parameters()]) if (len(st) and key in st[next(iter(st))]) else np.zeros(1)
        ga = b.grad_avg.value.copy() if b.grad_avg is not None else np.zeros(1)
        runs.append((theta, s1, ga, x, b.n_updates))
        e.close()
    (t0, s0, g0, x0, n0), (t1, s1, g1, x1, n1) = runs
    assert n0 == n1 == n_upd
    assert np.max(np.abs(t0 - t1)) < 1e-6 * np.max(np.abs(t0))
    assert s0.shape == s1.shape and np.max(np.abs(s0 - s1)) <= 1e-6 * max(np.max(np.abs(s0)), 1e-300)
    assert np.max(np.abs(g0 - g1)) <= 1e-9 * max(np.max(np.abs(g0)), 1e-300)
    assert np.max(np.abs(x0 - x1)) < 1e-4
    assert np.max(np.abs(t0 - np.concatenate([p.detach().numpy().ravel() for p in basis.NNBasis1D(-4, 4, hidden_layer_sizes=[48, 48, 48]).parameters()]))) > 1e-3


@pytest.mark.parametrize("opt,params,gavg", OPT_CASES, ids=["%s-%s-%s" % (o, "-".join(sorted(p)), g) for o, p, g in OPT_CASES])
def test_optimizer_step_kernel_matches_torch_optim(opt, params, gavg):
    """pyves_optimizer_step against the reference's optimiser objects (ves/bias.py:193-200) on the same synthetic sums for
    8 consecutive steps: gradient of the VES functional, running / exponentially weighted gradient average
    (pyves_b200.optim.RunningAverage on the host), then torch.optim.{SGD, RMSprop, Adam}.step().  1e-13 relative."""
    from pyves_b200.optim import RunningAverage
    rng = np.random.default_rng(12)
    P = 1000
    f64 = dict(dtype=torch.float64, device="cuda")
    theta0 = rng.standard_normal(P)
    prm = torch.nn.Parameter(torch.tensor(theta0))
    o = bias._make_torch_optimizer(opt, [prm], dict(params))
    kind, hyper, why = bias._torch_optimizer_plan(o)
    assert why is None and kind == opt
    ra = None if gavg is None else RunningAverage(None if gavg == "running" else gavg)
    theta, s1, s2, gs = torch.tensor(theta0, **f64), torch.zeros(P, **f64), torch.zeros(P, **f64), torch.zeros(P, **f64)
    for t in range(1, 9):
        sw = np.array([rng.uniform(10, 20)])
        sf = rng.standard_normal(P) * sw[0]
        ft = rng.standard_normal(P)
        g = ft - sf / sw[0]
        if ra is not None:
            g = ra.update(g)
        o.zero_grad()
        prm.grad = torch.tensor(g).clone()
        o.step()
        _lib.optimizer_step(0, kind, torch.tensor(sw, **f64), torch.tensor(sf, **f64), torch.tensor(ft, **f64),
                            None if gavg is None else ("running" if gavg == "running" else "ewma"), 0.0 if gavg in (None, "running") else gavg, t,
                            gs if gavg is not None else None, theta, s1, s2 if kind == "Adam" else None, hyper, t)
        ref = prm.detach().numpy()
        assert np.max(np.abs(theta.cpu().numpy() - ref)) < 1e-13 * np.max(np.abs(ref)), (t, np.max(np.abs(theta.cpu().numpy() - ref)))
    assert np.max(np.abs(ref - theta0)) > 1e-3


def test_device_updates_skip_an_empty_ensemble_average():
    """sum_w == 0 (no walker inside the averaging domain): both device optimiser kernels leave parameters and state as
    they are instead of dividing by zero (averaged_sgd_kernel had no guard)."""
    f64 = dict(dtype=torch.float64, device="cuda")
    K = 13
    alpha, abar = torch.randn(K, **f64), torch.randn(K, **f64)
    a_out, b_out = torch.empty(K, **f64), torch.empty(K, **f64)
    zero = torch.zeros(1, **f64)
    _lib.averaged_sgd_step(0, zero, torch.zeros(K, **f64), torch.zeros(K, K, **f64), torch.zeros(K, **f64), alpha, abar, None, 0.5, 0.4, 'ewma',
                           0.9, 1, a_out, b_out)
    assert torch.equal(a_out, alpha) and torch.equal(b_out, abar)
    theta, m, v, g = torch.randn(50, **f64), torch.randn(50, **f64), torch.rand(50, **f64), torch.randn(50, **f64)
    keep = [t.clone() for t in (theta, m, v, g)]
    _lib.optimizer_step(0, "Adam", zero, torch.randn(50, **f64), torch.randn(50, **f64), "running", 0.0, 3, g, theta, m, v, (1e-3, 0.9, 0.999, 1e-8), 3)
    for a, b in zip(keep, (theta, m, v, g)):
        assert torch.equal(a, b)
    _lib.optimizer_step(0, "Adam", torch.ones(1, **f64), torch.randn(50, **f64), torch.randn(50, **f64), "running", 0.0, 3, g, theta, m, v,
                        (1e-3, 0.9, 0.999, 1e-8), 3)
    assert not torch.equal(keep[0], theta) and torch.isfinite(theta).all()
    with pytest.raises(ValueError):
        _lib.optimizer_step(0, "Adam", zero, zero, zero, None, 0.0, 1, None, theta[:1].contiguous(), None, None, (1e-3, 0.9, 0.999, 1e-8), 1)


def test_host_parameter_upload_is_ordered_on_the_callers_stream():
    """pyves_set_bias_legendre2d / _mlp / _pathcv copy their parameters with cudaMemcpyAsync on the caller's stream
    (a synchronous copy on the legacy stream was not ordered against kernels on a non-blocking stream).
    A long launch followed at once by a re-upload and another launch, all on one side stream, equals the fully
    synchronised sequence."""
    rng = np.random.default_rng(4)
    w1, w2 = rng.standard_normal((32, 32)) * 0.05, rng.standard_normal((32, 32)) * 0.05
    mm = (-2.5, 2.5, -2.0, 2.0)
    res = []
    for synced in (True, False):
        e = _lib.Ensemble(200000, dims=2, seed=3, precision="fp32")
        e.set_integrator(1.0, KB * 300, 20.0, 0.01)
        e.set_potential(_lib.POT_DOUBLE_WELL_2D)
        assert e.init_positions_mc(mm, 1 / KT) == 0
        e.init_velocities()
        torch.cuda.synchronize()
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            e.set_bias_legendre2d(31, 31, mm, w1)
            e.step(300)                       # tens of milliseconds of work in flight ...
            if synced:
                torch.cuda.synchronize()
            e.set_bias_legendre2d(31, 31, mm, w2)     # ... while the next coefficients are uploaded
            e.step(50)
            x, _ = e.get_state()
        torch.cuda.synchronize()
        res.append(x.clone())
        e.close()
    assert torch.equal(res[0], res[1])


def _leg_bias():
    return bias.VESBias_Ensemble(basis.LegendreBasis1D(6, -2.5, 2.5, weights=np.zeros(7)), target.Target_WellTempered_x(201, 5.0), 1 / KT,
                                 "averaged_sgd", {"mu": 1.0, "second_order": True, "averaging": "running"}, target_update_every=3)


def _leg2_bias():
    return bias.VESBias_Ensemble(basis.LegendreBasis2D(7, 7, -2.5, 2.5, -2.0, 2.0), target.Target_WellTempered_2D(30, 30, 10.0), 1 / KT,
                                 "averaged_sgd", {"mu": 0.5, "averaging": "ewma", "decay": 0.8, "precondition": 1.0}, target_update_every=2)


def _nn_bias():
    torch.manual_seed(9)
    return bias.VESBias_Ensemble(basis.NNBasis1D(min=-2.5, max=2.5, axis='x', hidden_layer_sizes=[16, 16, 16], act=torch.nn.Tanh),
                                 target.Target_Uniform_HardSwitch_x(100), 1 / KT, "Adam", {"lr": 2e-3}, target_update_every=0, gradient_average=0.9)


@pytest.mark.parametrize("make,device_update,sample_every", [(_leg_bias, False, None), (_leg_bias, True, None), (_leg2_bias, True, None),
                                                             (_nn_bias, False, None), (_nn_bias, True, None), (_leg_bias, False, 20)],
                         ids=["leg1d-host", "leg1d-device", "leg2d-wt-device", "nn-adam-host", "nn-adam-device", "leg1d-host-sampled"])
def test_dynamic_run_resumes_bit_identical_from_checkpoint(make, device_update, sample_every, tmp_path):
    """a dynamic ensemble run interrupted at an arbitrary checkpoint and resumed from the pickled State
    (walkers, Philox counter, coefficients / NN parameters, optimiser iterates and averages, gradient average, target
    density, update count, pending moment samples) ends bit-identical (FP64) to the uninterrupted run -- both from the
    end-of-run checkpoint of a shorter run and from a mid-run checkpoint written right after an update."""
    pot = potentials.DoubleWellPotential2D()
    kw = dict(n_walkers=4096, seed=21, temp=300, precision="fp64", device_update=device_update, sample_every=sample_every)
    run = dict(chkevery=0, trajevery=0, energyevery=0, trajfile=None, energyfile=None)

    def fresh(**extra):
        sim = quiet(SingleParticleSimulation, pot, **kw, **extra)
        if "init_state" not in extra:
            sim.place_walkers(xmin=-2.5, xmax=2.5, ymin=-2, ymax=2)
        return sim

    def final(sim, b):
        st = sim.get_state()
        return st.positions.copy(), st.velocities.copy(), np.concatenate([p.detach().numpy().ravel() for p in b.model.parameters()]), b.n_updates

    # uninterrupted: 700 steps, updates at 100, 150, ..., 650
    bA = make()
    simA = fresh()
    simA.init_ves(bA, startafter=100, learnevery=50)
    quiet(simA, nsteps=700, chkfile=None, **run)
    ref = final(simA, bA)
    assert ref[3] == 12

    for cut, mid in ((330, False), (400, False), (400, True)):
        b1 = make()
        sim1 = fresh()
        sim1.init_ves(b1, startafter=100, learnevery=50)
        if mid:
            # a mid-run checkpoint (chkevery) is written right AFTER the update at its index; keep that first dump, as
            # a crash right after it would have
            dumps = []
            sim1._dump_state = lambda f, i, loop_step=None: dumps.append(pickle.dumps(sim1.get_state(loop_step=i if loop_step is None else loop_step)))
            quiet(sim1, nsteps=cut + 1, chkfile="unused", **dict(run, chkevery=cut))
            state = pickle.loads(dumps[0])
            assert state.loop_step == cut and state.last_update_step == cut
        else:
            # the end-of-run checkpoint of a shorter run: written BEFORE the bookkeeping of loop index `cut`
            chk = str(tmp_path / ("chk_%d.pkl" % cut))
            quiet(sim1, nsteps=cut, chkfile=chk, **run)
            with open(chk, "rb") as fh:
                state = pickle.load(fh)
            assert state.loop_step == cut and state.last_update_step == (cut // 50) * 50 - (50 if cut % 50 == 0 else 0)
        assert state.bias_state is not None and state.step == cut
        b2 = make()
        sim2 = fresh(init_state=state)
        sim2.init_ves(b2, startafter=100, learnevery=50)
        quiet(sim2, nsteps=700 - cut, chkfile=None, **run)
        got = final(sim2, b2)
        assert got[3] == ref[3]
        assert np.array_equal(got[2], ref[2]), (cut, mid, np.max(np.abs(got[2] - ref[2])))
        assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1]), (cut, mid)


def test_device_update_archive_is_optional(tmp_path):
    """device_update=True writes the reference's per-update archive model_loc.iter{n} (ves/bias.py:249-253) only when
    archive_updates=True; the files reload with plain torch.jit."""
    pot = potentials.DoubleWellPotential2D()
    for archive in (False, True):
        loc = str(tmp_path / ("m%d.pt" % archive))
        b = bias.VESBias_Ensemble(basis.LegendreBasis1D(4, -2.5, 2.5, weights=np.zeros(5)), target.Target_Uniform_HardSwitch_x(101), 1 / KT,
                                  "averaged_sgd", {"mu": 1.0}, model_loc=loc)
        sim = quiet(SingleParticleSimulation, pot, n_walkers=2048, seed=1, device_update=True, archive_updates=archive)
        sim.place_walkers(xmin=-2.5, xmax=2.5, ymin=-2, ymax=2)
        sim.init_ves(b, startafter=10, learnevery=10)
        quiet(sim, nsteps=45, chkevery=0, trajevery=0, energyevery=0, chkfile=None, trajfile=None, energyfile=None)
        names = sorted(p.name for p in tmp_path.iterdir() if p.name.startswith("m%d.pt" % archive))
        if archive:
            assert names == ["m1.pt", "m1.pt.iter1", "m1.pt.iter2", "m1.pt.iter3", "m1.pt.iter4"]
            mod = torch.jit.load(str(tmp_path / "m1.pt.iter4"))
            assert torch.allclose(mod.weights.detach(), b.model.weights.detach())
        else:
            assert names == ["m0.pt"]


def test_native_communicator_single_rank():
    """pyves_comm_* / pyves_allreduce_moments (the collective of the path in the C ABI, NCCL bound at run time): a
    one-rank communicator is created from a unique id, the in-place all-reduce is the identity, bad arguments raise.
    The multi-rank case runs in bench.py (--comm native is the default; SCALE_r02 covers 2 / 4 / 8 ranks)."""
    assert _lib.Communicator.nccl_version() >= 20000
    uid = _lib.Communicator.unique_id()
    assert len(uid) == 128
    c = _lib.Communicator(0, 1, 0, uid)
    t = torch.arange(257, dtype=torch.float64, device="cuda") * 0.5
    ref = t.clone()
    c.allreduce_(t)
    torch.cuda.synchronize()
    assert torch.equal(t, ref)
    with pytest.raises(ValueError):
        c.allreduce_(t.float())
    with pytest.raises(ValueError):
        _lib.Communicator(0, 2, 5, uid)
    c.close()
