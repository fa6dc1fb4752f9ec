import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from pyves_b200 import _lib, basis, bias, potentials, target
from pyves_b200.bias import DeviceUpdater
potentials._Potential.quiet = True
KT = 8.3145e-3 * 300; KB = 8.31446261815324e-3
def ens():
    e = _lib.Ensemble(20000, dims=2, seed=7, precision="fp32")
    e.set_integrator(1.0, KB * 300, 20.0, 0.01)
    e.set_potential(_lib.POT_SZABO_BEREZHKOVSKII, (2.2, 4.0, 4.04, 4.84))
    assert e.init_positions_mc((-4, 4, -4, 4), 1 / KT, ntrials=2000) == 0
    e.init_velocities(); return e
def mk():
    torch.manual_seed(2)
    m = basis.NNBasis1D(min=-4, max=4, axis='x', hidden_layer_sizes=[48, 48, 48])
    return bias.VESBias_Ensemble(m, target.Target_Uniform_HardSwitch_x(200), 1 / KT, "Adam", {"lr": 1e-3}, target_update_every=0)
bh, bd = mk(), mk()
eh, ed = ens(), ens()
bh.force.apply(eh)
upd = DeviceUpdater(bd, ed)
th = lambda b: np.concatenate([p.detach().numpy().ravel() for p in b.model.parameters()])
for it in range(26):
    eh.step(20); ed.step(20)
    swh, sfh, _ = eh.moments()
    bufs = upd.moment_buffers(); ed.moments(out=bufs)
    ft_h = bh.target_average(); ft_d = upd.f_target.cpu().numpy()
    gh = ft_h - (sfh / swh).cpu().numpy(); gd = ft_d - (bufs[1] / bufs[0]).cpu().numpy()
    bh.update_from_moments(swh, sfh); bh.force.apply(eh)
    upd.update(*bufs)
    tdev = upd.theta.cpu().numpy()
    d = np.abs(th(bh) - tdev); k = int(np.argmax(d))
    xh, _ = eh.get_state(host=True); xd, _ = ed.get_state(host=True)
    print(it, 'max dtheta', d.max(), 'at', k, 'g host/dev', gh[k], gd[k], 'max dg', np.abs(gh - gd).max(), 'dft', np.abs(ft_h - ft_d).max(), 'dx', np.abs(xh - xd).max(),
          'n(|g|<1e-7)', int((np.abs(gh) < 1e-7).sum()), 'n(g==0)', int((gh == 0).sum()))
