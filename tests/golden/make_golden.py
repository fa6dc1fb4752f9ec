#!/usr/bin/env python
"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_golden.py

It imports ``/root/reference/ves`` with import stubs for the packages that are absent from the
image (openmm, openmmtorch, matplotlib), evaluates the reference's own modules in float64 and
writes ``tests/golden/*.json``.  Nothing here is read at test time except the JSON files.
"""
import contextlib
import io
import json
import os
import sys
import tempfile
import types
from unittest import mock

import numpy as np

REF = os.environ.get("PYVES_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))


def install_stubs():
    class CustomExternalForce:            # openmm.openmm.CustomExternalForce
        def __init__(self, expr):
            self.expression = expr

        def addParticle(self, *a):
            return 0

    omm = types.ModuleType("openmm")
    omm_inner = types.ModuleType("openmm.openmm")
    omm_inner.CustomExternalForce = CustomExternalForce
    omm_inner.State = object
    omm.openmm = omm_inner
    omm.unit = mock.MagicMock()
    sys.modules["openmm"] = omm
    sys.modules["openmm.openmm"] = omm_inner
    sys.modules["openmm.unit"] = omm.unit
    ot = types.ModuleType("openmmtorch")
    ot.TorchForce = lambda path: ("TorchForce", path)
    sys.modules["openmmtorch"] = ot
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.animation", "matplotlib.cm"):
        sys.modules[name] = mock.MagicMock()
    plt = sys.modules["matplotlib.pyplot"]
    plt.subplots.return_value = (mock.MagicMock(), mock.MagicMock())
    plt.subplots.return_value[1].get_ylim.return_value = (0.0, 1.0)
    sys.modules["matplotlib"].pyplot = plt


def dump(name, obj):
    with open(os.path.join(OUT, name), "w") as fh:
        json.dump(obj, fh, indent=None, separators=(",", ":"))
    print("wrote", name, os.path.getsize(os.path.join(OUT, name)), "bytes")


def lst(a):
    return np.asarray(a, dtype=np.float64).tolist()


def main():
    install_stubs()
    sys.path.insert(0, REF)
    import torch
    import warnings
    warnings.filterwarnings("ignore")
    from ves.basis import LegendreBasis1D, NNBasis1D, LegendreBasis2DPathCV
    from ves.bias import VESBias_SingleParticle_1D, HarmonicBias_SingleParticle_1D_ForceModule
    from ves.target import Target_Uniform_HardSwitch_x
    import ves.potentials as P
    import ves.config_creation as CC
    from ves.visualization import VisualizePotential2D

    def energy_force(module, pos, per_row=False):
        """Reference evaluation path: TorchScript module forward + autograd (ves/bias.py:52,58)."""
        scripted = torch.jit.script(module)
        pos = np.asarray(pos, dtype=np.float64)
        if not per_row:
            t = torch.tensor(pos, requires_grad=True)
            e = scripted(t)
            e.sum().backward()
            g = np.zeros_like(pos) if t.grad is None else t.grad.numpy()   # degree 0: no dependence
            return e.detach().numpy().reshape(-1), -g
        E, F = [], []
        for row in pos:
            t = torch.tensor(row.reshape(1, -1), requires_grad=True)
            e = scripted(t)
            e.sum().backward()
            E.append(float(e.detach().numpy().reshape(-1)[0]))
            F.append(-t.grad.numpy()[0])
        return np.array(E), np.array(F)

    rng = np.random.default_rng(20221019)

    # ------------------------------------------------------------------ Legendre 1D
    w_b1 = np.array([-4.48286631e+00, -1.74999996e+00, -5.65477074e+00, 7.26710048e-05, -8.92828224e+00])
    cases = []

    def leg_case(tag, degree, lo, hi, axis, w, pos):
        m = LegendreBasis1D(degree, lo, hi, axis=axis, weights=np.asarray(w, dtype=np.float64))
        E, F = energy_force(m, pos)
        cases.append(dict(tag=tag, degree=degree, min=lo, max=hi, axis=axis, weights=lst(w),
                          positions=lst(pos), E=lst(E), F=lst(F)))

    xs = np.array([-3.0, -2.5, -1.4, 0.0, 0.3, 1.4, 2.5, 2.6])
    leg_case("B1", 4, -2.5, 2.5, 'x', w_b1, np.stack([xs, 0.7 * np.ones_like(xs), np.zeros_like(xs)], 1))
    ys = np.array([-4.5, -1.0, 0.25, 3.3, 5.999])
    leg_case("B2", 12, -4.0, 6.0, 'y', [(-1.0) ** n / (n + 1) for n in range(13)],
             np.stack([0.1 * np.ones_like(ys), ys, np.zeros_like(ys)], 1))
    for tag, deg, lo, hi, axis in [("cfg1", 5, -2.5, 2.5, 'x'), ("cfg3x", 12, -3.0, 6.0, 'x'),
                                   ("cfg3y", 5, -4.0, 6.0, 'y'), ("d0", 0, -1.0, 1.0, 'x'),
                                   ("d1", 1, -1.0, 2.0, 'y'), ("d15", 15, -2.0, 2.0, 'x'),
                                   ("d63", 63, -2.5, 2.5, 'x')]:
        w = rng.standard_normal(deg + 1) * 3.0
        pos = np.stack([rng.uniform(lo - 0.5, hi + 0.5, 64), rng.uniform(lo - 0.5, hi + 0.5, 64),
                        rng.standard_normal(64) * 0.03], 1)
        leg_case(tag, deg, lo, hi, axis, w, pos)
    dump("legendre1d.json", cases)

    # ------------------------------------------------------------------ NN basis
    cases = []
    for tag, sizes, lo, hi, axis, act in [("cfg4", [48, 48, 48], -4.0, 4.0, 'x', "relu"),
                                         ("example", [128, 128], -4.0, 4.0, 'x', "relu"),
                                         ("one", [16], -1.0, 3.0, 'y', "relu"),
                                         ("tanh", [24, 24], -2.0, 2.0, 'x', "tanh")]:
        torch.manual_seed(7 + len(cases))
        m = NNBasis1D(lo, hi, axis=axis, hidden_layer_sizes=sizes,
                      act=torch.nn.ReLU if act == "relu" else torch.nn.Tanh)
        with torch.no_grad():                      # widen the spread so ReLUs switch inside [lo, hi]
            for p in m.parameters():
                p.mul_(2.5)
        theta = np.concatenate([p.detach().numpy().ravel() for p in m.parameters()])
        pos = np.stack([rng.uniform(lo - 1, hi + 1, 48), rng.uniform(lo - 1, hi + 1, 48),
                        np.zeros(48)], 1)
        E, F = energy_force(m, pos, per_row=True)   # the reference module only handles N = 1
        cases.append(dict(tag=tag, sizes=sizes, min=lo, max=hi, axis=axis, act=act, theta=lst(theta),
                          positions=lst(pos), E=lst(E), F=lst(F)))
    dump("nnbasis1d.json", cases)

    # ------------------------------------------------------------------ path CV + harmonic
    cases = []
    for tag, deg, xi, yi, lam, w, phi, pos in [
        ("B3", 4, np.linspace(-4, 5, 20), np.linspace(-4, 6, 20), 100.0, w_b1, 2.0,
         np.array([[0.3, -0.2, 0.0], [1.0, 1.5, 0.0]])),
        ("slip", 8, np.linspace(-3.5, 5.5, 100), np.linspace(-3.5, 5.5, 100) + 0.3 * np.sin(np.linspace(0, 3, 100)),
         100.0, rng.standard_normal(9), 5.0,
         np.stack([rng.uniform(-4, 6, 32), rng.uniform(-4, 6, 32), np.zeros(32)], 1)),
    ]:
        m = LegendreBasis2DPathCV(deg, np.asarray(xi), np.asarray(yi), lam, weights=np.asarray(w, dtype=np.float64), phi=phi)
        E, F = energy_force(m, pos)
        cases.append(dict(tag=tag, degree=deg, x_i=lst(xi), y_i=lst(yi), lam=lam, weights=lst(w), phi=phi,
                          positions=lst(pos), E=lst(E), F=lst(F)))
    dump("pathcv.json", cases)

    hm = HarmonicBias_SingleParticle_1D_ForceModule(12.5, 0.4, axis='y')
    pos = np.array([[0.1, -0.3, 0.0]])
    E, F = energy_force(hm, pos)
    dump("harmonic.json", dict(k=12.5, s0=0.4, axis='y', positions=lst(pos), E=lst(E), F=lst(F)))

    # ------------------------------------------------------------------ update() known answers
    cases = []
    beta = 1 / (8.3145e-3 * 300)
    traj = np.zeros((200, 3))
    traj[:, 0] = -1.4 + 0.1 * np.random.default_rng(0).standard_normal(200)

    def update_case(tag, make_model, opt, opt_params, mesh, update_steps, traj, nupdates=1):
        model = make_model()
        theta0 = np.concatenate([p.detach().numpy().ravel() for p in model.parameters()])
        tmp = tempfile.mkdtemp()
        target = Target_Uniform_HardSwitch_x(mesh)
        bias = VESBias_SingleParticle_1D(model, target, beta, opt, opt_params,
                                         os.path.join(tmp, "model.pt"), update_steps=update_steps)
        losses, grads, thetas = [], [], []
        for u in range(nupdates):
            buf = io.StringIO()
            with contextlib.redirect_stdout(buf):
                bias.update(traj[: len(traj) - (nupdates - 1 - u) * 20])
            losses.append(float(buf.getvalue().split("Loss =")[1].split()[0]))
            grads.append(np.concatenate([p.grad.numpy().ravel() for p in model.parameters()]).tolist())
            thetas.append(np.concatenate([p.detach().numpy().ravel() for p in model.parameters()]).tolist())
        files = sorted(os.listdir(tmp))
        cases.append(dict(tag=tag, optimizer=opt, optimizer_params=opt_params, mesh=mesh, beta=beta,
                          update_steps=update_steps, traj=lst(traj), nupdates=nupdates, theta0=lst(theta0),
                          loss_printed=losses, grad=grads, theta=thetas, files=files))

    def leg0():
        return LegendreBasis1D(5, -2.5, 2.5, weights=np.zeros(6))

    def leg_rand():
        return LegendreBasis1D(5, -2.5, 2.5, weights=np.array([0.3, -1.2, 0.8, 0.05, -0.4, 0.2]))

    update_case("B4_sgd", leg0, "SGD", {"lr": 0.1}, 11, 100, traj)
    update_case("leg_adam3", leg_rand, "Adam", {"lr": 0.01}, 21, 150, traj, nupdates=3)
    update_case("leg_rmsprop2", leg_rand, "RMSprop", {"lr": 0.01}, 21, 5000, traj, nupdates=2)

    def nn_small():
        torch.manual_seed(3)
        return NNBasis1D(-4.0, 4.0, hidden_layer_sizes=[8, 8])

    traj2 = np.zeros((120, 3))
    traj2[:, 0] = 2.2 + 0.3 * np.random.default_rng(1).standard_normal(120)
    traj2[:, 1] = 2.2 + 0.3 * np.random.default_rng(2).standard_normal(120)
    update_case("nn_adam2", nn_small, "Adam", {"lr": 0.001}, 20, 5000, traj2, nupdates=2)
    dump("update.json", cases)

    # ------------------------------------------------------------------ potentials, target, placement
    pots = {}
    with contextlib.redirect_stdout(io.StringIO()):
        objs = dict(SingleWellPotential2D=P.SingleWellPotential2D(), DoubleWellPotential2D=P.DoubleWellPotential2D(),
                    SlipBondPotential2D=P.SlipBondPotential2D(),
                    SlipBondPotential2D_forced=P.SlipBondPotential2D(force_x=1.5, force_y=-0.5, y_0=0.8, y_scale=4.0,
                                                                   y_shift=3.0, xy_0=0.2, xy_scale=2.5),
                    CatchBondPotential2D=P.CatchBondPotential2D(), SzaboBerezhkovskiiPotential=P.SzaboBerezhkovskiiPotential(),
                    MullerBrownPotential=P.MullerBrownPotential(),
                    SingleWellPotential1D=P.SingleWellPotential1D(), DoubleWellPotential1D=P.DoubleWellPotential1D())
    x = rng.uniform(-3, 3, 40)
    y = rng.uniform(-3, 3, 40)
    for name, o in objs.items():
        if name.endswith("1D"):
            pots[name] = dict(x=lst(x), U=lst(o.potential(x)), expr=o.force)
        else:
            pots[name] = dict(x=lst(x), y=lst(y), U=lst(o.potential(x, y)), expr=o.force)
    dump("potentials.json", pots)

    t = Target_Uniform_HardSwitch_x(11)
    dump("target.json", dict(mesh=11, x=lst(t.x), p=lst(t.p)))

    np.random.seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        c1 = CC.singleParticle2D_init_coord(objs["DoubleWellPotential2D"], 300, xmin=-2.5, xmax=2.5, ymin=-2, ymax=2)
        c2 = CC.singleParticle2D_init_coord(objs["SlipBondPotential2D"], 300, xmin=-8, xmax=10, ymin=-6, ymax=8)
    dump("placement.json", dict(numpy_seed=0, double_well=lst(c1), slip_bond_after=lst(c2)))

    # ------------------------------------------------------------------ analytic FES projections
    fes = {}
    for name, xr, yr in [("DoubleWellPotential2D", [-2.25, 2.25], [-2, 2]),
                         ("SlipBondPotential2D", [-8, 10], [-6, 8]),
                         ("SzaboBerezhkovskiiPotential", [-4, 4], [-4, 4])]:
        vis = VisualizePotential2D(objs[name], temp=300, xrange=xr, yrange=yr)
        _, _, gx, Fx = vis.plot_projection_x()
        _, _, gy, Fy = vis.plot_projection_y()
        fes[name] = dict(temp=300, xrange=xr, yrange=yr, mesh=200, x=lst(gx[::8]), Fx=lst(Fx[::8]),
                         y=lst(gy[::8]), Fy=lst(Fy[::8]))
    dump("fes.json", fes)


if __name__ == "__main__":
    main()
