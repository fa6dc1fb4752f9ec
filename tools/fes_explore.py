#!/usr/bin/env python
"""Convergence experiments behind the FES tests of tests/test_gpu_fes.py (GPU box):

    python tools/fes_explore.py cfg0|cfg2|cfg3 [key=value ...]

Prints one JSON line per run: the FES RMSE (kT) against the analytic projection and the settings used.
"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyves_b200 import _lib, basis, bias, fes, potentials, target  # noqa: E402

potentials._Potential.quiet = True
KB, KB_PY, TEMP = 8.31446261815324e-3, 8.3145e-3, 300.0
KT = KB_PY * TEMP
BETA = 1.0 / KT


def ensemble(n, pot, box, seed, precision="fp32", ntrials=2000):
    e = _lib.Ensemble(n, dims=2, seed=seed, precision=precision)
    e.set_integrator(1.0, KB * TEMP, 20.0, 0.01)
    e.set_potential(pot.kind, pot.params)
    assert e.init_positions_mc(box, BETA, ntrials=ntrials) == 0
    e.init_velocities()
    return e


def cfg2(n=65536, updates=1200, stride=200, mu=3.0, decay=0.9, seed=3, second=1, averaging="ewma", lo=-3.0, hi=6.0, domain="box", device=1):
    pot = potentials.SlipBondPotential2D()
    model = basis.LegendreBasis1D(12, lo, hi, 'x', weights=np.zeros(13))
    b = bias.VESBias_Ensemble(model, target.Target_Uniform_HardSwitch_x(801), BETA, "averaged_sgd",
                              {"mu": mu, "second_order": bool(second), "averaging": averaging, "decay": decay}, sample_domain=domain)
    e = ensemble(n, pot, (-8, 10, -6, 8), seed)
    x, Fx = fes.projection_x(pot, TEMP, [lo, hi], [-6, 8], mesh=200)
    trace = []
    upd = bias.DeviceUpdater(b, e) if device else None
    if not device:
        b.force.apply(e)
    t0 = time.time()
    for u in range(updates):
        e.step(stride)
        if device:
            bufs = upd.moment_buffers(bool(second))
            e.moments(cov=bool(second), out=bufs)
            upd.update(*bufs)
        else:
            sw, sf, sff = e.moments(cov=bool(second))
            b.update_from_moments(sw, sf, sff)
            b.force.apply(e)
        if (u + 1) % max(1, updates // 8) == 0:
            if device:
                upd.sync_to_host()
            V = model(torch.tensor(np.stack([x, 0 * x], 1), device="cuda")).detach().cpu().numpy()
            trace.append(round(fes.rmse_kt((-V - (-V).min()) / KT, Fx, clip=20.0), 4))
    torch.cuda.synchronize()
    pos, _ = e.get_state(host=True)
    inside = float(np.mean((pos[0] >= lo) & (pos[0] <= hi)))
    return dict(cfg="cfg2", n=n, updates=updates, stride=stride, mu=mu, decay=decay, second=second, averaging=averaging, lo=lo, hi=hi,
                domain=domain, rmse_trace=trace, rmse=trace[-1], frac_inside=inside, seconds=time.time() - t0)


def cfg3(n=100000, updates=1000, stride=100, lr=1e-3, gavg="none", seed=5, opt="Adam", act="relu", yr=12.0, tau=0.0, vavg=0.98, domain="box",
         init_seed=0, hidden="48,48,48"):
    """NN bias through the device-resident update loop; lr_t = lr / (1 + t / tau) when tau > 0."""
    pot = potentials.SzaboBerezhkovskiiPotential()
    torch.manual_seed(init_seed)
    model = basis.NNBasis1D(min=-4, max=4, axis='x', hidden_layer_sizes=[int(v) for v in str(hidden).split(",")],
                            act=torch.nn.ReLU if act == "relu" else torch.nn.Tanh)
    ga = None if gavg in ("none", "None") else (gavg if gavg == "running" else float(gavg))
    b = bias.VESBias_Ensemble(model, target.Target_Uniform_HardSwitch_x(200), BETA, opt, {"lr": lr}, target_update_every=0, gradient_average=ga,
                              sample_domain=domain)
    e = ensemble(n, pot, (-4, 4, -4, 4), seed)
    upd = bias.DeviceUpdater(b, e)
    x, Fx = fes.projection_x(pot, TEMP, [-4, 4], [-yr, yr], mesh=200)
    trace = []
    t0 = time.time()
    xg = torch.tensor(np.stack([x, 0 * x], 1), device="cuda")
    Vbar = None
    for u in range(updates):
        e.step(stride)
        if tau > 0:
            upd.lr = lr / (1.0 + u / tau)
        bufs = upd.moment_buffers()
        e.moments(out=bufs)
        upd.update(*bufs)
        Vn, _ = upd.ev.eval_bias(xg, want_force=False)
        Vn = Vn - Vn.mean()
        Vbar = Vn.clone() if Vbar is None else vavg * Vbar + (1 - vavg) * Vn
        if (u + 1) % max(1, updates // 10) == 0:
            V = Vn.cpu().numpy()
            F = (-V - (-V).min()) / KT
            Vb = Vbar.cpu().numpy()
            Fb = (-Vb - (-Vb).min()) / KT
            inner = np.abs(x) <= 3.5
            trace.append((round(fes.rmse_kt(F, Fx, clip=20.0), 3), round(fes.rmse_kt(F[inner], Fx[inner], clip=20.0), 3), round(fes.rmse_kt(Fb, Fx, clip=20.0), 3)))
    torch.cuda.synchronize()
    pos, _ = e.get_state(host=True)
    return dict(cfg="cfg3", n=n, updates=updates, stride=stride, lr=lr, gavg=gavg, opt=opt, act=act, tau=tau, domain=domain, seed=seed, init_seed=init_seed,
                hidden=hidden, rmse_trace=trace, rmse=trace[-1][0], rmse_vavg=trace[-1][2], frac_inside=float(np.mean(np.abs(pos[0]) <= 4)),
                seconds=time.time() - t0)


def cfg0(n=4096, nsteps=200000, every=10, nbins=100, seed=11, fit="legfit"):
    pot = potentials.DoubleWellPotential2D()
    xs, Fx = fes.projection_x(pot, TEMP, [-2.25, 2.25], [-2, 2])
    w = -np.polynomial.legendre.legfit(xs / 2.5, KT * Fx, 5)
    model = basis.LegendreBasis1D(5, -2.5, 2.5, 'x', weights=w)
    e = ensemble(n, pot, (-2.5, 2.5, -2, 2), seed, ntrials=100)
    bias.StaticBias_SingleParticle(model, None).force.apply(e)
    counts = torch.zeros(nbins, dtype=torch.float64, device="cuda")
    t0 = time.time()
    e.step(2000)
    for _ in range(nsteps // every):
        e.step(every)
        pos, _ = e.get_state()
        V, _ = e.eval_bias(pos.T.contiguous(), want_force=False)
        e.histogram((-2.25, 2.25), nbins, axis=0, row_weights=torch.exp(BETA * V).contiguous(), counts=counts)
    torch.cuda.synchronize()
    c = counts.cpu().numpy()
    centres = -2.25 + 4.5 * (np.arange(nbins) + 0.5) / nbins
    bF = fes.counts_to_fes(c)
    F_ref = np.interp(centres, xs, Fx)
    return dict(cfg="cfg0", n=n, nsteps=nsteps, every=every, rmse=fes.rmse_kt(bF, F_ref, clip=20.0), seconds=time.time() - t0,
                max_err=float(np.max(np.abs((bF - bF.min()) - (F_ref - F_ref.min()))[F_ref <= 20])))


if __name__ == "__main__":
    fn = globals()[sys.argv[1]]
    kw = {}
    for a in sys.argv[2:]:
        k, v = a.split("=")
        try:
            kw[k] = int(v)
        except ValueError:
            try:
                kw[k] = float(v)
            except ValueError:
                kw[k] = v
    print(json.dumps(fn(**kw)), flush=True)
