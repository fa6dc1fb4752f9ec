import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from pyves_b200 import _lib
KB = 8.31446261815324e-3; MM = (-2.5, 2.5, -2.0, 2.0)
def ens(n, K, w, variant, prec="fp32"):
    e = _lib.Ensemble(n, dims=2, seed=1, precision=prec)
    e.set_integrator(1.0, KB * 300, 20.0, 0.01); e.set_potential(_lib.POT_DOUBLE_WELL_2D)
    e.set_bias_legendre2d(K - 1, K - 1, MM, w); e.set_tuning(variant); return e
rng = np.random.default_rng(0)
K = int(sys.argv[1]) if len(sys.argv) > 1 else 64
w = rng.standard_normal((K, K)) * 3.0 / (1.0 + np.add.outer(np.arange(K), np.arange(K)))
n, nsteps = 5000, 60
x0 = np.stack([rng.uniform(-2.7, 2.7, n), rng.uniform(-2.2, 2.2, n)])
v0 = rng.standard_normal((2, n)) * 1.58
noise = rng.standard_normal((nsteps, 2, n))
runs = {}
for name, variant, prec in (("cc32", -1, "fp32"), ("tc32", 7, "fp32"), ("ref64", -1, "fp64")):
    dt = torch.float32 if prec == "fp32" else torch.float64
    e = ens(n, K, w, variant, prec)
    e.set_state(x0.astype(e.np_dtype), v0.astype(e.np_dtype))
    tn = torch.tensor(noise, dtype=dt, device="cuda")
    xs = []
    for s in range(nsteps):
        e.step(1, tn[s:s + 1].contiguous())
        xs.append(e.get_state(host=True)[0].astype(np.float64).copy())
    runs[name] = np.array(xs)
    e.close()
for s in (0, 1, 2, 5, 10, 20, 40, 59):
    d_cc = np.abs(runs["cc32"][s] - runs["ref64"][s]); d_tc = np.abs(runs["tc32"][s] - runs["ref64"][s])
    print(s, "cc32-ref64 max %.3e  tc32-ref64 max %.3e  median %.2e %.2e  argmax walker x0 %s" % (d_cc.max(), d_tc.max(), np.median(d_cc), np.median(d_tc), x0[:, np.argmax(d_tc.max(0))]))
# single launch of 60 steps vs 60 launches
for name, variant in (("cc32", -1), ("tc32", 7)):
    e = ens(n, K, w, variant)
    e.set_state(x0.astype(np.float32), v0.astype(np.float32))
    e.step(nsteps, torch.tensor(noise, dtype=torch.float32, device="cuda"))
    x = e.get_state(host=True)[0].astype(np.float64)
    print(name, "one launch vs 60 launches max", np.abs(x - runs[name][-1]).max(), " vs ref64", np.abs(x - runs["ref64"][-1]).max())
    e.close()
