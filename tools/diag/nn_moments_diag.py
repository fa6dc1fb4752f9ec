#!/usr/bin/env python
"""Diagnostic: per-parameter error of the NN parameter-gradient moments (per-row kernel vs through-the-pieces) against the FP64 oracle."""
import os, sys, zlib
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import biasspec
from oracle import ves, potentials, mlp


def rng_spec(kind, rng, sizes):
    return dict(kind="mlp", axis='x', min=-4.0, max=4.0, sizes=sizes, act=0, theta=rng.standard_normal(mlp.n_params(sizes)) * 0.4)
sizes = [48, 48, 48]
rng = np.random.default_rng(zlib.crc32(str(sizes).encode()))
spec = rng_spec("mlp", rng, sizes=sizes)
n = 20011
pos = np.stack([rng.uniform(-0.2, 1.2, n) * (spec["max"] - spec["min"]) + spec["min"], rng.uniform(-2.5, 2.5, n)], 1)
fr = biasspec.oracle_basis(spec, pos)
sw_r, sf_r, _ = ves.moments(fr, np.ones(n))
scale = np.maximum(np.abs(fr).T @ np.ones(n), 1.0)
for prec, dt in (("fp32", torch.float32), ("fp64", torch.float64)):
    e = biasspec.make_ensemble(n, 2, 0, prec, potentials.DOUBLE_WELL_2D, None, spec)
    e.set_state(torch.tensor(pos, dtype=dt, device="cuda").T.contiguous(), None)
    for variant in (8, 6):
        e.set_tuning(variant)
        sw, sf, _ = e.moments()
        err = np.abs(sf.cpu().numpy() - sf_r) / scale
        k = int(np.argmax(err))
        print(prec, "variant", variant, "max err %.3e at parameter %d of %d (value %.4e, oracle %.4e, scale %.3e); median %.2e; n>1e-4: %d"
              % (err[k], k, len(err), sf.cpu().numpy()[k], sf_r[k], scale[k], np.median(err), int((err > 1e-4).sum())))
    e.close()
