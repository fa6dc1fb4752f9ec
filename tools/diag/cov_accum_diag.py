import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from pyves_b200 import _lib
rng = np.random.default_rng(3)
for n in (100000, 1000000, 10000000):
    e = _lib.Ensemble(n, dims=2, seed=1, precision="fp32")
    e.set_integrator(1.0, 2.494, 20.0, 0.01); e.set_potential(_lib.POT_DOUBLE_WELL_2D)
    e.set_bias_legendre2d(15, 15, (-2.5, 2.5, -2, 2), np.zeros(256))
    g = torch.Generator(device="cuda").manual_seed(5)
    pos = torch.empty((2, n), dtype=torch.float32, device="cuda")
    pos[0].uniform_(-2.5, 2.5, generator=g); pos[1].uniform_(-2.0, 2.0, generator=g)
    e.set_state(pos, None)
    e.set_tuning(98); _, _, ref = e.moments(cov=True); e.set_tuning(97)
    ref = ref.cpu().numpy(); scale = np.sqrt(np.outer(np.diag(ref), np.diag(ref)))
    for lg in (12, 13, 14, 16):
        e.set_tuning(1000 + lg)
        _, _, c = e.moments(cov=True); c = c.cpu().numpy()
        d = (c - ref) / scale
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); e.moments(cov=True); b.record(); torch.cuda.synchronize()
        print(json.dumps(dict(n=n, rows_per_cta=2 ** lg, max_err=float(np.abs(d).max()), diag_mean_signed=float(np.mean(np.diag(d))), diag_max=float(np.abs(np.diag(d)).max()),
                              offdiag_max=float(np.abs(d - np.diag(np.diag(d))).max()), ms=a.elapsed_time(b))), flush=True)
    e.set_tuning(1013)
    e.close()
