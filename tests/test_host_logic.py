"""CPU tests: host-side mirror of the reference API, the C-ABI surface, optimisers, sharding."""
import contextlib
import io
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def test_abi_header_binding_and_library_agree():
    """Every symbol include/pyves_b200.h declares is bound by _lib.SIGNATURES and exported by the .so."""
    from pyves_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "pyves_b200.h")).read()
    declared = set(re.findall(r"\b(pyves_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    if not os.path.exists(_lib.LIB_PATH):
        pytest.skip("library not built")
    lib = _lib.load()                       # dlopen only; no compute call without a GPU
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.pyves_abi_version() == 2
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (pyves_[a-z0-9_]+)", out))
    assert exported == declared, exported ^ declared


def test_product_has_no_cpu_fallback():
    """Nothing under pyves_b200/ imports the oracle, and constructing an ensemble without CUDA raises."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "pyves_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f
    import torch
    if not torch.cuda.is_available():
        from pyves_b200 import _lib
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            _lib.Ensemble(4)
        from pyves_b200.basis import LegendreBasis1D
        with pytest.raises(RuntimeError):
            LegendreBasis1D(3, -1, 1)(torch.zeros(2, 3, dtype=torch.float64))


def test_potentials_match_reference_golden(golden):
    from pyves_b200 import potentials as P
    g = golden("potentials.json")
    objs = dict(SingleWellPotential2D=quiet(P.SingleWellPotential2D), DoubleWellPotential2D=quiet(P.DoubleWellPotential2D),
                SlipBondPotential2D=quiet(P.SlipBondPotential2D),
                SlipBondPotential2D_forced=quiet(P.SlipBondPotential2D, force_x=1.5, force_y=-0.5, y_0=0.8, y_scale=4.0,
                                                 y_shift=3.0, xy_0=0.2, xy_scale=2.5),
                CatchBondPotential2D=quiet(P.CatchBondPotential2D), SzaboBerezhkovskiiPotential=quiet(P.SzaboBerezhkovskiiPotential),
                MullerBrownPotential=quiet(P.MullerBrownPotential), SingleWellPotential1D=quiet(P.SingleWellPotential1D),
                DoubleWellPotential1D=quiet(P.DoubleWellPotential1D))
    for name, c in g.items():
        o = objs[name]
        assert o.force == c["expr"], name            # the Lepton expression string is part of the API
        x = np.array(c["x"])
        U = o.potential(x) if name.endswith("1D") else o.potential(x, np.array(c["y"]))
        assert np.allclose(U, c["U"], rtol=1e-12, atol=1e-12), name
    assert objs["SlipBondPotential2D"].params == (0.0, 0.0, 1.0, 5.0, 4.0, 0.0, 2.0)
    assert objs["SzaboBerezhkovskiiPotential"].params == (2.2, 4.0, 4.04, 4.840000000000001)
    with pytest.raises(NotImplementedError):
        P.Potential2D.potential(objs["SlipBondPotential2D"].__class__.__mro__[1], 0, 0)
    # the expression is printed at construction, as upstream
    buf = io.StringIO()
    P._Potential.quiet = False
    with contextlib.redirect_stdout(buf):
        P.DoubleWellPotential2D()
    assert "[Potential] Initializing potential with expression:" in buf.getvalue()


def test_targets(golden):
    from pyves_b200 import target as T
    g = golden("target.json")
    t = T.Target_Uniform_HardSwitch_x(g["mesh"])
    assert np.array_equal(t.x, g["x"]) and np.array_equal(t.p, g["p"])
    assert abs(t.quadrature_weights().sum() - 1) < 1e-15 and t.update(np.zeros(g["mesh"]), 0.4) is False
    # well-tempered: with V = -(1 - 1/gamma) F the fixed point is p ~ exp(-(beta/gamma) F)
    beta, gamma = 0.4, 10.0
    wt = T.Target_WellTempered_x(201, gamma)
    F = 30 * wt.x ** 2
    for _ in range(200):
        wt.update(-(1 - 1 / gamma) * F, beta)
    ref = np.exp(-(beta / gamma) * F)
    assert np.allclose(wt.p / wt.p.max(), ref / ref.max(), atol=1e-6)
    assert abs(np.trapezoid(wt.p, wt.x) - 1) < 1e-12
    w2 = T.Target_WellTempered_2D(20, 30, gamma)
    assert w2.x.shape == (600, 2) and abs(w2.x[0, 0] + 0.95) < 1e-12 and abs(w2.x[1, 1] - (-1 + 3 / 30)) < 1e-12
    F2 = 10 * (w2.x[:, 0] ** 2 + w2.x[:, 1] ** 2)
    for _ in range(200):
        w2.update(-(1 - 1 / gamma) * F2, beta)
    r2 = np.exp(-(beta / gamma) * F2)
    assert np.allclose(w2.p / w2.p.max(), r2 / r2.max(), atol=1e-6)
    assert abs(w2.p.sum() * 4 / 600 - 1) < 1e-12
    with pytest.raises(ValueError):
        T.Target_WellTempered_x(10, 1.0)


def test_placement_matches_reference_golden(golden):
    from pyves_b200 import config_creation as CC, potentials as P
    g = golden("placement.json")
    np.random.seed(g["numpy_seed"])
    c1 = quiet(CC.singleParticle2D_init_coord, quiet(P.DoubleWellPotential2D), 300, xmin=-2.5, xmax=2.5, ymin=-2, ymax=2)
    c2 = quiet(CC.singleParticle2D_init_coord, quiet(P.SlipBondPotential2D), 300, xmin=-8, xmax=10, ymin=-6, ymax=8)
    assert c1.shape == (1, 3) and np.array_equal(c1, g["double_well"]) and np.array_equal(c2, g["slip_bond_after"])
    with pytest.raises(RuntimeError, match="Could not place particle"):
        quiet(CC.singleParticle2D_init_coord, quiet(P.DoubleWellPotential2D), 300, ntrials=3, xmin=2.4, xmax=2.5, ymin=1.9, ymax=2)
    c = quiet(CC.singleParticle1D_init_coord, quiet(P.SingleWellPotential1D), 300, xmin=-0.5, xmax=0.5)
    assert c.shape == (1, 3) and c[0, 1] == 0


def test_optimisers_match_oracle():
    from oracle import ves
    from pyves_b200 import optim
    rng = np.random.default_rng(0)
    f = rng.standard_normal((500, 7))
    sw, sf, sff = ves.moments(f, rng.uniform(0.5, 1.5, 500))
    ft = rng.standard_normal(7)
    g1, h1 = ves.gradient_hessian(sw, sf, sff, ft, 0.4)
    g2, h2 = optim.gradient_hessian(sw, sf, sff, ft, 0.4)
    assert np.array_equal(g1, g2) and np.array_equal(h1, h2)
    a, b = ves.AveragedSGD(np.ones(7), 0.3), optim.AveragedSGD(np.ones(7), 0.3, second_order=True, averaging='running')
    for _ in range(5):
        assert np.allclose(a.step(g1, h1), b.step(g2, h2), rtol=0, atol=1e-15)
    a, b = ves.AveragedSGD(np.ones(7), 0.3, decay=0.8), optim.AveragedSGD(np.ones(7), 0.3, averaging='ewma', decay=0.8)
    for _ in range(5):
        assert np.allclose(a.step(g1), b.step(g2), rtol=0, atol=1e-15)
    st = b.state_dict()
    c = optim.AveragedSGD(np.zeros(7))
    c.load_state_dict(st)
    assert np.array_equal(c.step(g2, h2), b.step(g2, h2))
    ra = optim.RunningAverage()
    for k in range(1, 5):
        v = ra.update(np.full(3, float(k)))
    assert np.allclose(v, 2.5)


def test_host_side_validation_messages():
    from pyves_b200 import basis, bias, langevin_dynamics as ld
    import torch
    with pytest.raises(ValueError, match="NNBasis needs at least one hidden layer"):
        basis.NNBasis1D(0, 1)
    with pytest.raises(NotImplementedError):
        basis.LegendreBasis2DRadialCV(3, 0, 0, 1, 1)
    m = basis.LegendreBasis1D(5, -2.5, 2.5, axis='q', weights=np.zeros(6))
    with pytest.raises(ValueError, match="Invalid axis"):
        m(torch.zeros(1, 3))
    assert tuple(basis.LegendreBasis1D(5, -1, 1).weights.shape) == (6,)
    assert basis.LegendreBasis1D(5, -1, 1).weights.dtype == torch.float64
    nn = basis.NNBasis1D(-4, 4, hidden_layer_sizes=[48, 48, 48])
    assert sum(p.numel() for p in nn.parameters()) == 4849 and len(nn.layers) == 6
    with pytest.raises(NotImplementedError, match="This bias cannot be updated"):
        bias.Bias().update(None)
    with pytest.raises(ValueError, match="Invalid axis"):
        bias.HarmonicBias_SingleParticle_1D(1.0, 0.0, axis='z', model_loc=None)
    with pytest.raises(ValueError, match="Requested optimizer not yet supported."):
        bias.VESBias_SingleParticle_1D(basis.LegendreBasis1D(2, -1, 1), None, 0.4, "LBFGS", {}, None)
    sim = ld.SingleParticleSimulation.__new__(ld.SingleParticleSimulation)
    with pytest.raises(ValueError, match="Cannot start VES updates before timestep 2."):
        sim.init_ves(bias.Bias(), startafter=1)
    sim.init_ves(bias.Bias(), static=True)
    assert sim.biased and sim.static and sim.startafter is None


def test_ensemble_bias_host_logic_without_gpu():
    """VESBias_Ensemble pieces that need no device: sample_domain validation and its hand-over through the bias
    descriptor, DeviceUpdater.supports, and the path-CV target averages (node sums of Legendre polynomials)."""
    from pyves_b200 import basis, bias, target
    kT = 8.3145e-3 * 300
    sgd = {"mu": 0.5, "averaging": "ewma", "decay": 0.8}
    m2 = basis.LegendreBasis2D(3, 3, -1, 1, -1, 1)
    with pytest.raises(ValueError, match="sample_domain"):
        bias.VESBias_Ensemble(m2, target.Target_Uniform_2D(8, 8), 1 / kT, "averaged_sgd", sgd, sample_domain="plane")
    for dom, flag in (("box", True), ("all", False)):
        b = bias.VESBias_Ensemble(m2, target.Target_Uniform_2D(8, 8), 1 / kT, "averaged_sgd", sgd, sample_domain=dom)
        d = b.force
        assert d.kind == "legendre2d" and d.moment_inside is flag
        assert bias.DeviceUpdater.supports(b)
    assert bias.StaticBias_SingleParticle(m2, None).force.moment_inside is None      # static biases leave the domain alone
    # NN bases: SGD / RMSprop / Adam in their plain form run on the device; options the kernel lacks stay on the host
    nn = basis.NNBasis1D(-4, 4, hidden_layer_sizes=[8, 8])
    tx = target.Target_Uniform_HardSwitch_x(50)
    for opt, prm in (("Adam", {"lr": 1e-3}), ("SGD", {"lr": 1e-2, "momentum": 0.5}), ("RMSprop", {"lr": 1e-3}),
                     ("Adam", {"lr": 1e-3, "betas": (0.8, 0.99), "eps": 1e-6})):
        assert bias.DeviceUpdater.supports(bias.VESBias_Ensemble(nn, tx, 1 / kT, opt, prm, gradient_average='running'))
    for opt, prm in (("Adam", {"lr": 1e-3, "amsgrad": True}), ("Adam", {"lr": 1e-3, "weight_decay": 1e-2}), ("SGD", {"lr": 1e-2, "nesterov": True, "momentum": 0.5}),
                     ("RMSprop", {"lr": 1e-3, "centered": True})):
        assert bias.DeviceUpdater.why_not(bias.VESBias_Ensemble(nn, tx, 1 / kT, opt, prm)) is not None
    assert not bias.DeviceUpdater.supports(bias.VESBias_Ensemble(basis.NNBasis1D(-4, 4, hidden_layer_sizes=[8]), tx, 1 / kT, "Adam", {"lr": 1e-3}))
    assert not bias.DeviceUpdater.supports(bias.VESBias_Ensemble(m2, target.Target_Uniform_2D(8, 8), 1 / kT, "averaged_sgd", sgd,
                                                                 gradient_average='running'))
    # path CV: uniform target on [-1, 1] -> <P_k>_p = delta_k0 (trapezoid rule on 2001 nodes); well-tempered refresh
    x_i, y_i = np.linspace(-2, 2, 30), np.zeros(30)
    mp = basis.LegendreBasis2DPathCV(6, x_i, y_i, 100.0, weights=np.array([0.0, 1.0, -0.5, 0.0, 0.2, 0.0, 0.0]))
    bp = bias.VESBias_Ensemble(mp, target.Target_Uniform_HardSwitch_x(2001), 1 / kT, "averaged_sgd", sgd, target_update_every=0)
    ft = bp.target_average()
    assert abs(ft[0] - 1) < 1e-12 and np.max(np.abs(ft[1:])) < 1e-5
    assert not bias.DeviceUpdater.supports(bp)
    bw = bias.VESBias_Ensemble(mp, target.Target_WellTempered_x(401, 4.0), 1 / kT, "averaged_sgd", sgd, target_update_every=1)
    s = np.asarray(bw.target.x)
    V = np.polynomial.legendre.legval(s, mp.weights.detach().numpy())
    bw.refresh_target()
    p_ref = np.exp(-(1 / kT / 4.0) * ((-V - np.log(0.5) * kT) - (-V - np.log(0.5) * kT).min()))
    wq = np.ones_like(s); wq[0] = wq[-1] = 0.5
    p_ref /= np.sum(p_ref * wq) * (s[1] - s[0])
    assert np.max(np.abs(bw.target.p - p_ref)) < 1e-12
    F = bw.fes()
    assert F.min() == 0 and F.shape == s.shape


def test_torchscript_exports_are_scriptable_and_match_oracle(tmp_path, golden):
    """model_loc files (ves/bias.py:52-53) load with plain torch.jit and reproduce the golden values."""
    import torch
    from pyves_b200 import basis, bias
    c = [x for x in golden("legendre1d.json") if x["tag"] == "B1"][0]
    m = basis.LegendreBasis1D(c["degree"], c["min"], c["max"], axis=c["axis"], weights=np.array(c["weights"]))
    b = bias.StaticBias_SingleParticle(m, str(tmp_path / "model.pt"))
    loaded = torch.jit.load(str(tmp_path / "model.pt"))
    pos = torch.tensor(c["positions"], dtype=torch.float64, requires_grad=True)
    E = loaded(pos)
    E.sum().backward()
    assert np.allclose(E.detach().numpy(), c["E"], rtol=1e-12) and np.allclose(-pos.grad.numpy(), c["F"], rtol=1e-10, atol=1e-12)
    assert b.force.kind == "legendre1d" and b.force.params["degree"] == 4
    c = [x for x in golden("nnbasis1d.json") if x["tag"] == "cfg4"][0]
    nn = basis.NNBasis1D(c["min"], c["max"], axis=c["axis"], hidden_layer_sizes=c["sizes"])
    o = 0
    with torch.no_grad():
        for p in nn.parameters():
            p.copy_(torch.tensor(c["theta"][o:o + p.numel()], dtype=torch.float64).reshape(p.shape))
            o += p.numel()
    E = nn.export_torchscript()(torch.tensor(c["positions"], dtype=torch.float64))
    assert np.allclose(E.detach().numpy(), c["E"], rtol=1e-10)
    c = golden("pathcv.json")[0]
    pc = basis.LegendreBasis2DPathCV(c["degree"], np.array(c["x_i"]), np.array(c["y_i"]), c["lam"], np.array(c["weights"]), c["phi"])
    assert np.allclose(pc.export_torchscript()(torch.tensor(c["positions"], dtype=torch.float64)).detach().numpy(), c["E"], rtol=1e-9)
    l2 = basis.LegendreBasis2D(3, 4, -1, 2, -2, 2, weights=np.arange(20.0))
    from oracle import legendre
    p2 = np.random.default_rng(0).uniform(-2, 2, (10, 2))
    assert np.allclose(l2.export_torchscript()(torch.tensor(p2)).detach().numpy(), legendre.legendre2d(p2, 3, 4, -1, 2, -2, 2, np.arange(20.0))[0])


def test_shard_covers_all_walkers():
    from pyves_b200.parallel import shard
    for n, w in [(10, 3), (1000000, 8), (7, 7), (5, 1)]:
        parts = [shard(n, r, w) for r in range(w)]
        assert parts[0][0] == 0 and parts[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(parts[:-1], parts[1:]))
        assert max(hi - lo for lo, hi in parts) - min(hi - lo for lo, hi in parts) <= 1


def test_trajectory_reader(tmp_path):
    from pyves_b200.utils import TrajectoryReader
    p = tmp_path / "traj.dat"
    with open(p, "w") as fh:
        fh.write("# t[ps]    x [nm]    y [nm]    z[nm]\n")
        for i in range(10):
            fh.write("{:10.5f}\t{:10.7f}\t{:10.7f}\t{:10.7f}\n".format(i * 0.01, i, -i, 0.0))
    t, xyz = TrajectoryReader(str(p)).read_traj()
    assert t.shape == (10,) and xyz.shape == (10, 3) and xyz[3, 1] == -3
    t2, _ = TrajectoryReader(str(p)).read_traj(skip=3)
    assert np.allclose(t2, [0, 0.03, 0.06, 0.09])


WORKER = r"""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, %(root)r)
from oracle import legendre, ves
from pyves_b200 import optim
from pyves_b200.parallel import reduce_moments, shard
rank, world = int(sys.argv[1]), 2
os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=sys.argv[2], RANK=str(rank), WORLD_SIZE=str(world))
dist.init_process_group("gloo", rank=rank, world_size=world)
pos = np.random.default_rng(0).uniform(-2.5, 2.5, (1001, 2))
lo, hi = shard(1001, rank, world)
f = legendre.basis2d(pos[lo:hi], 3, 3, -2.5, 2.5, -2, 2)
sw, sf, sff = ves.moments(f)
out = reduce_moments(torch.tensor([sw]), torch.tensor(sf), torch.tensor(sff))
ft = np.zeros(16); ft[0] = 1
g, h = optim.gradient_hessian(float(out[0]), out[1].numpy(), out[2].numpy(), ft, 0.4)
opt = optim.AveragedSGD(np.zeros(16), 0.5, second_order=True, averaging='running')
np.save(sys.argv[3] + "/rank%%d.npy" %% rank, opt.step(g, h))
dist.destroy_process_group()
"""


def test_two_rank_moment_reduction_gloo(tmp_path):
    """world_size 2 on CPU (gloo): sharded moments + one all_reduce give the single-process update on every rank."""
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker.py"
    script.write_text(WORKER % dict(root=ROOT))
    procs = [subprocess.Popen([sys.executable, str(script), str(r), str(port), str(tmp_path)]) for r in range(2)]
    for p in procs:
        assert p.wait(timeout=180) == 0
    from oracle import legendre, ves
    from pyves_b200 import optim
    pos = np.random.default_rng(0).uniform(-2.5, 2.5, (1001, 2))
    sw, sf, sff = ves.moments(legendre.basis2d(pos, 3, 3, -2.5, 2.5, -2, 2))
    ft = np.zeros(16)
    ft[0] = 1
    g, h = optim.gradient_hessian(sw, sf, sff, ft, 0.4)
    ref = optim.AveragedSGD(np.zeros(16), 0.5, second_order=True, averaging='running').step(g, h)
    r0, r1 = np.load(tmp_path / "rank0.npy"), np.load(tmp_path / "rank1.npy")
    assert np.array_equal(r0, r1) and np.allclose(r0, ref, rtol=1e-12, atol=1e-14)


def test_histogram_fes_semantics():
    """ves/visualization.py:475-478: empty bins take the smallest non-zero count, beta F = -log counts, shifted to
    min 0; rmse_kt compares min-shifted profiles where the reference is below the clip."""
    from pyves_b200 import fes
    c = np.array([0.0, 4.0, 16.0, 0.0, 2.0])
    bF = fes.counts_to_fes(c)
    assert np.allclose(bF, -np.log([2.0, 4.0, 16.0, 2.0, 2.0]) + np.log(16.0))
    assert bF.min() == 0.0 and c[0] == 0.0                     # the input is not modified
    with pytest.raises(ValueError):
        fes.counts_to_fes(np.zeros(4))
    ref = np.array([0.0, 1.0, 5.0, 30.0])
    est = ref + 7.0                                             # a constant offset is not an error
    est[3] = 0.0                                                # ... and bins above the clip are ignored
    assert fes.rmse_kt(est, ref, clip=20.0) < 1e-12
    assert abs(fes.rmse_kt(ref + np.array([0.0, 0.3, 0.0, 0.0]), ref, clip=20.0) - 0.3 / np.sqrt(3)) < 1e-12


def test_update_state_checkpoint_round_trip_without_gpu():
    """bias.state_dict() / load_state_dict(): averaged-SGD iterates and averages, torch.optim state, gradient running
    average, target density and update count survive a pickle round trip into a freshly constructed bias (SURVEY 8(f3);
    the reference archives only the TorchScript module, ves/bias.py:249-253)."""
    import pickle
    import torch
    from pyves_b200 import basis, bias, target
    from pyves_b200.langevin_dynamics import State
    kT = 8.3145e-3 * 300
    rng = np.random.default_rng(1)

    def make_leg():
        return bias.VESBias_Ensemble(basis.LegendreBasis1D(6, -2.5, 2.5, weights=np.zeros(7)), target.Target_WellTempered_x(101, 5.0), 1 / kT,
                                     "averaged_sgd", {"mu": 0.7, "second_order": True, "averaging": "running"})
    a = make_leg()
    a._f_target = rng.standard_normal(7)              # stands in for the device quadrature (no GPU here)
    for _ in range(5):
        g = rng.standard_normal(7)
        a.optimizer.step(g, np.eye(7) * 0.1)
    a.n_updates = 5
    a.target._p = rng.uniform(0.2, 0.8, 101)
    st = State(np.zeros((4, 2)), np.zeros((4, 2)), 1234, 7, loop_step=600, last_update_step=600, bias_state=a.state_dict(),
               pending=[np.ones(1), np.arange(7.0), None])
    st2 = pickle.loads(pickle.dumps(st))
    assert st2.loop_step == 600 and st2.last_update_step == 600 and st2.step == 1234 and st2.pending[2] is None
    b = make_leg()
    b.load_state_dict(st2.bias_state)
    assert np.array_equal(b.optimizer.alpha, a.optimizer.alpha) and np.array_equal(b.optimizer.abar, a.optimizer.abar)
    assert b.optimizer.n == 5 and b.n_updates == 5 and np.array_equal(b.target.p, a.target.p) and np.array_equal(b._f_target, a._f_target)
    g = rng.standard_normal(7)
    assert np.array_equal(a.optimizer.step(g, np.eye(7)), b.optimizer.step(g, np.eye(7)))      # continues identically
    with pytest.raises(ValueError, match="optimizer"):
        bias.VESBias_Ensemble(basis.NNBasis1D(-4, 4, hidden_layer_sizes=[8, 8]), target.Target_Uniform_HardSwitch_x(50), 1 / kT, "Adam",
                              {"lr": 1e-3}).load_state_dict(st2.bias_state)

    def make_nn():
        torch.manual_seed(3)
        return bias.VESBias_Ensemble(basis.NNBasis1D(-4, 4, hidden_layer_sizes=[8, 8]), target.Target_Uniform_HardSwitch_x(50), 1 / kT, "Adam",
                                     {"lr": 1e-3}, gradient_average=0.9)
    c = make_nn()
    for _ in range(3):
        grad = c.grad_avg.update(rng.standard_normal(105))
        c.optimizer.zero_grad()
        bias._set_grads(c.model, grad)
        c.optimizer.step()
    c.n_updates = 3
    d = make_nn()
    d.load_state_dict(pickle.loads(pickle.dumps(c.state_dict())))
    g = rng.standard_normal(105)
    for x in (c, d):
        x.optimizer.zero_grad()
        bias._set_grads(x.model, x.grad_avg.update(g))
        x.optimizer.step()
    for p, q in zip(c.model.parameters(), d.model.parameters()):
        assert torch.equal(p, q)
    assert d.grad_avg.n == 4


def test_empty_ensemble_average_skips_the_update():
    """sum_w == 0 (every walker outside the basis box with sample_domain='box'): the host update changes nothing and
    warns, the same rule the device kernels apply (averaged_sgd_kernel / optimizer_step_kernel)."""
    from pyves_b200 import basis, bias, optim, target
    kT = 8.3145e-3 * 300
    b = bias.VESBias_Ensemble(basis.LegendreBasis1D(4, -1, 1, weights=np.array([0.0, 1.0, 0.5, 0.0, -0.2])), target.Target_Uniform_HardSwitch_x(11),
                              1 / kT, "averaged_sgd", {"mu": 0.5})
    b._f_target = np.array([1.0, 0, 0, 0, 0])
    before = (b.optimizer.alpha.copy(), b.optimizer.abar.copy())
    with pytest.warns(UserWarning, match="VES update skipped"):
        b.update_from_moments(np.zeros(1), np.zeros(5), None)
    assert np.array_equal(b.optimizer.alpha, before[0]) and np.array_equal(b.optimizer.abar, before[1]) and b.n_updates == 1
    with pytest.raises(optim.EmptyEnsembleAverage):
        optim.gradient_hessian(0.0, np.zeros(5), None, np.zeros(5), 1 / kT)


def test_trajectory_reader_formats(tmp_path):
    """TrajectoryReader: 'txyz' -> (t, xyz), 'xyz' -> xyz only (ves/utils.py:37-40), anything else raises before reading."""
    from pyves_b200.utils import TrajectoryReader
    f = tmp_path / "t.dat"
    f.write_text("# t x y z\n0.0 1 2 3\n0.1 4 5 6\n0.2 7 8 9\n")
    t, xyz = TrajectoryReader(str(f)).read_traj()
    assert t.tolist() == [0.0, 0.1, 0.2] and xyz.shape == (3, 3) and xyz[2].tolist() == [7, 8, 9]
    g = tmp_path / "x.dat"
    g.write_text("# x y z\n1 2 3\n4 5 6\n")
    assert TrajectoryReader(str(g), format='xyz').read_traj().tolist() == [[1, 2, 3], [4, 5, 6]]
    assert TrajectoryReader(str(f)).read_traj(skip=2)[1].shape == (2, 3)
    with pytest.raises(ValueError, match="Invalid format"):
        TrajectoryReader(str(tmp_path / "missing.dat"), format='tx').read_traj()


def test_relu_nnbasis1d_is_piecewise_linear_in_its_input():
    """The two facts the piecewise-constant NN paths of the CUDA side rest on (csrc/bias.cuh BiasPwl1D, csrc/moments.cu pwl_hist_kernel),
    checked on the CPU against the network itself (oracle/mlp.py): dV/ds looked up in the compiled table equals the derivative of the
    network, and parameter-gradient sums over rows equal the sums over one pseudo row per piece."""
    from oracle import mlp, pwl
    rng = np.random.default_rng(12)
    for sizes in ([16], [7, 9], [48, 48, 48], [10, 12, 8, 6]):
        theta = rng.standard_normal(mlp.n_params(sizes)) * 0.5
        brk, slope = pwl.compile_pwl(theta, sizes)
        assert np.all(np.diff(brk) >= 0) and len(slope) == len(brk) + 1
        s = rng.uniform(-1.5, 2.5, 4000)
        if len(brk):                                   # stay clear of the switch points (the derivative jumps there)
            s = s[np.min(np.abs(s[:, None] - brk[None, :]), axis=1) > 1e-9]
        _, dV = mlp.forward(s, theta, sizes, mlp.ACT_RELU)
        scale = max(1.0, float(np.max(np.abs(dV))))
        assert np.max(np.abs(pwl.slope_at(s, brk, slope) - dV)) < 1e-10 * scale, sizes
        w = rng.uniform(0, 2, s.size)
        ref = mlp.param_grad(s, theta, sizes, mlp.ACT_RELU, w)
        got, rows = pwl.param_grad_through_pieces(s, theta, sizes, w)
        assert rows <= len(brk) + 1
        assert np.max(np.abs(got - ref)) < 1e-9 * max(1.0, float(np.max(np.abs(ref)))), sizes


def test_folded_legendre_derivatives_and_monic_recursion():
    """The two rewrites of the tensor-core Legendre step kernel (csrc/langevin_tc.cu), checked on the CPU against the recursion of
    oracle/legendre.py: forces from Legendre VALUES only through the folded matrices WA / WB, and the scaled monic recursion."""
    from oracle import legendre, legendre_fold as lf
    rng = np.random.default_rng(3)
    s = np.concatenate([rng.uniform(-1, 1, 500), [-1.0, 1.0, 0.0]])
    cc, d, kappa = lf.monic_table(64)
    assert set(np.unique(d)) <= {0, 1} and kappa.min() > 0.70 and kappa.max() < 1.42
    P, _ = legendre.legendre_table(s, 63)
    q = lf.monic_values(s, 64)
    assert np.max(np.abs(q)) < 1.42 and np.max(np.abs(kappa[:, None] * q - P)) < 1e-12
    for kx, ky in ((16, 16), (64, 64), (5, 11), (17, 3), (1, 50), (2, 1)):
        w = rng.standard_normal((kx, ky))
        sx, sy = rng.uniform(-1, 1, 300), rng.uniform(-1, 1, 300)
        Px, dPx = legendre.legendre_table(sx, kx - 1)
        Py, dPy = legendre.legendre_table(sy, ky - 1)
        gx_ref = np.sum(dPx * (w @ Py), axis=0)
        gy_ref = np.sum(Px * (w @ dPy), axis=0)
        gx, gy = lf.gradient_folded(sx, sy, w)
        scale = max(1.0, float(np.max(np.abs(gx_ref))), float(np.max(np.abs(gy_ref))))
        assert np.max(np.abs(gx - gx_ref)) < 1e-10 * scale and np.max(np.abs(gy - gy_ref)) < 1e-10 * scale, (kx, ky)
