#!/usr/bin/env python
"""Host-side cost of one VES iteration of bench.py (where the time between two step launches goes)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from pyves_b200.parallel import reduce_moments
args = bench.parse_args()
beta = 1.0 / (bench.KB_PY * bench.TEMP)
from pyves_b200.potentials import _Potential
_Potential.quiet = True
bias = bench.make_bias(args, beta)
ens = bench.make_ensemble(args, 0, 0)
bias.force.apply(ens)
T = {k: 0.0 for k in ("step", "moments", "reduce+d2h", "update", "descriptor", "apply")}
def tick(): torch.cuda.synchronize(); return time.perf_counter()
N = 50
for it in range(N + 5):
    if it == 5: T = {k: 0.0 for k in T}
    t0 = tick(); ens.step(args.stride)
    t1 = tick(); m = ens.moments()
    t2 = tick(); sw, sf, _ = reduce_moments(*m); sw = sw.cpu(); sf = sf.cpu()
    t3 = tick(); bias.update_from_moments(sw, sf)
    t4 = tick(); d = bias.force
    t5 = tick(); d.apply(ens)
    t6 = tick()
    for k, dt in zip(T, (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t6 - t5)): T[k] += dt
for k, v in T.items(): print("%-12s %.3f ms" % (k, v / N * 1e3))
