"""Test helper: one bias description applied to both the CUDA handle and the oracle."""
import numpy as np

from oracle import legendre, mlp, pathcv, potentials
from pyves_b200 import _lib

AX = {'x': 0, 'y': 1}


def apply(ens, spec):
    k = spec["kind"]
    if k == "none":
        ens.set_bias_none()
    elif k == "leg1d":
        ens.set_bias_legendre1d(AX[spec["axis"]], spec["degree"], spec["min"], spec["max"], spec["weights"])
    elif k == "leg2d":
        ens.set_bias_legendre2d(spec["deg_x"], spec["deg_y"], spec["minmax"], spec["weights"])
    elif k == "mlp":
        ens.set_bias_mlp(AX[spec["axis"]], spec["min"], spec["max"], spec["sizes"], spec.get("act", 0), spec["theta"])
    elif k == "pathcv":
        ens.set_bias_pathcv(spec["degree"], spec["x_i"], spec["y_i"], spec["lam"], spec["weights"], spec["phi"])
    elif k == "harmonic":
        ens.set_bias_harmonic(AX[spec["axis"]], spec["k"], spec["s0"])
    else:
        raise KeyError(k)


def oracle_energy_force(spec, pos):
    """pos [N, D] -> (E [N], F [N, D])"""
    pos = np.asarray(pos, dtype=np.float64)
    k = spec["kind"]
    if k == "none":
        return np.zeros(len(pos)), np.zeros_like(pos)
    if k == "leg1d":
        return legendre.legendre1d(pos, spec["degree"], spec["min"], spec["max"], spec["axis"], spec["weights"])
    if k == "leg2d":
        m = spec["minmax"]
        return legendre.legendre2d(pos, spec["deg_x"], spec["deg_y"], m[0], m[1], m[2], m[3], spec["weights"])
    if k == "mlp":
        return mlp.nnbasis1d(pos, spec["min"], spec["max"], spec["axis"], spec["theta"], spec["sizes"], spec.get("act", 0))
    if k == "pathcv":
        return pathcv.legendre_pathcv(pos, spec["degree"], spec["x_i"], spec["y_i"], spec["lam"], spec["weights"], spec["phi"])
    if k == "harmonic":
        return pathcv.harmonic(pos, spec["k"], spec["s0"], spec["axis"])
    raise KeyError(k)


def oracle_basis(spec, pos):
    """f [N, K]: basis values (Legendre) or parameter gradients (MLP)."""
    pos = np.asarray(pos, dtype=np.float64)
    k = spec["kind"]
    if k == "leg1d":
        return legendre.basis1d(pos, spec["degree"], spec["min"], spec["max"], spec["axis"])
    if k == "leg2d":
        m = spec["minmax"]
        return legendre.basis2d(pos, spec["deg_x"], spec["deg_y"], m[0], m[1], m[2], m[3])
    if k == "pathcv":
        s, _, _, _ = pathcv.path_cv(pos, spec["x_i"], spec["y_i"], spec["lam"])
        P, _ = legendre.legendre_table(np.clip(2 * s - 1, -1, 1), spec["degree"])
        return P.T.copy()
    if k == "mlp":
        s = (pos[:, AX[spec["axis"]]] - spec["min"]) / (spec["max"] - spec["min"])
        return np.stack([mlp.param_grad(s[i:i + 1], spec["theta"], spec["sizes"], spec.get("act", 0)) for i in range(len(s))])
    raise KeyError(k)


def oracle_inside(spec, pos):
    """Rows whose scaled coordinate lies inside the basis interval on every axis the basis uses
    (the rows pyves_set_moment_domain(1) keeps)."""
    pos = np.asarray(pos, dtype=np.float64)
    k = spec["kind"]
    if k == "leg1d":
        return legendre.scale_clamp(pos[:, AX[spec["axis"]]], spec["min"], spec["max"])[1].astype(bool)
    if k == "leg2d":
        m = spec["minmax"]
        return (legendre.scale_clamp(pos[:, 0], m[0], m[1])[1].astype(bool)
                & legendre.scale_clamp(pos[:, 1], m[2], m[3])[1].astype(bool))
    if k == "pathcv":
        s, _, _, _ = pathcv.path_cv(pos, spec["x_i"], spec["y_i"], spec["lam"])
        return (2 * s - 1 >= -1) & (2 * s - 1 <= 1)
    if k == "mlp":                                   # NNBasis1D: scaled coordinate in [0, 1] (the support of the target)
        s = (pos[:, AX[spec["axis"]]] - spec["min"]) / (spec["max"] - spec["min"])
        return (s >= 0) & (s <= 1)
    raise KeyError(k)


def total_force_fn(pot_kind, pot_params, spec):
    def f(x):
        return potentials.force(pot_kind, x, pot_params) + oracle_energy_force(spec, x)[1]
    return f


def make_ensemble(n, dims, seed, precision, pot_kind, pot_params, spec, consts=None, scheme="openmm_langevin"):
    from oracle import langevin
    k = consts or langevin.constants()
    ens = _lib.Ensemble(n, dims=dims, seed=seed, precision=precision)
    ens.set_integrator(k["mass"], k["kT"], k["friction"], k["dt"], scheme)
    ens.set_potential(pot_kind, pot_params if pot_params is not None else potentials.default_params(pot_kind))
    apply(ens, spec)
    return ens
