"""GPU parity tests: the CUDA library, called through its C ABI, against the CPU oracle.

Tolerances are the north-star's: energies/forces 1e-6 relative in FP64 and 1e-4 in FP32;
trajectories driven by an identical injected noise stream 1e-6 relative over 10^3 steps (FP64).
The integrator, the 2D tensor-product basis and the ensemble moments have no runnable reference
(parity unpinned, see oracle/__init__.py); the 1D Legendre, NN and path-CV evaluations are also
checked directly against the golden fixtures produced by the unmodified reference.
"""
import zlib

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import langevin, legendre, mlp, philox, potentials  # noqa: E402
from tests import biasspec  # noqa: E402
from pyves_b200 import _lib  # noqa: E402

DEV = "cuda:0"
PREC = [("fp64", 1e-6, torch.float64), ("fp32", 1e-4, torch.float32)]


def rel_err(a, b, floor=1.0):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor)))


def rng_spec(kind, rng, **kw):
    if kind == "leg1d":
        d = kw.get("degree", 5)
        return dict(kind="leg1d", axis=kw.get("axis", 'x'), degree=d, min=kw.get("min", -2.5), max=kw.get("max", 2.5),
                    weights=rng.standard_normal(d + 1) * kw.get("scale", 2.0))
    if kind == "leg2d":
        dx, dy = kw["deg_x"], kw["deg_y"]
        decay = 1.0 / (1.0 + np.add.outer(np.arange(dx + 1), np.arange(dy + 1)))
        return dict(kind="leg2d", deg_x=dx, deg_y=dy, minmax=kw.get("minmax", (-2.5, 2.5, -2.0, 2.0)),
                    weights=rng.standard_normal((dx + 1, dy + 1)) * decay * kw.get("scale", 3.0))
    if kind == "mlp":
        sizes = kw["sizes"]
        return dict(kind="mlp", axis=kw.get("axis", 'x'), min=kw.get("min", -4.0), max=kw.get("max", 4.0), sizes=sizes,
                    act=kw.get("act", 0), theta=rng.standard_normal(mlp.n_params(sizes)) * kw.get("scale", 0.4))
    raise KeyError(kind)


# ---------------------------------------------------------------------------------------------------
def test_abi_version_and_errors():
    lib = _lib.load()
    assert lib.pyves_abi_version() == 2
    with pytest.raises(ValueError):
        _lib.Ensemble(0)
    with pytest.raises(ValueError):
        _lib.Ensemble(4, dims=5)
    e = _lib.Ensemble(4)
    with pytest.raises(RuntimeError):          # stepping before configuration
        e.step(1)
    with pytest.raises(ValueError, match="Invalid axis"):
        e.set_bias_legendre1d(2, 3, -1, 1, np.zeros(4))
    with pytest.raises(ValueError, match="at least one hidden layer"):
        e.set_bias_mlp(0, 0, 1, [], 0, np.zeros(0))
    with pytest.raises(ValueError):
        e.set_potential(42)
    with pytest.raises(ValueError):
        e.set_potential(_lib.POT_SLIP_BOND_2D, [1.0, 2.0])
    e.close()


def test_velocity_and_placement_streams_match_oracle():
    k = langevin.constants()
    for prec, tol, _ in PREC:
        e = biasspec.make_ensemble(5000, 3, 1234, prec, potentials.DOUBLE_WELL_2D, None, dict(kind="none"))
        e.set_walker_offset(10 ** 10 + 7)       # exercises the high counter word
        e.init_velocities()
        _, vel = e.get_state(host=True)
        ref = langevin.maxwell_boltzmann(1234, 10 ** 10 + 7, 5000, k, 3)
        assert rel_err(vel.T, ref) < (1e-12 if prec == "fp64" else 2e-5)
        beta = 1 / (langevin.KB_PY * 300)
        nf = e.init_positions_mc((-2.5, 2.5, -2.0, 2.0), beta)
        assert nf == 0
        pos, _ = e.get_state(host=True)
        en = lambda x, y: potentials.energy_grad_xy(potentials.DOUBLE_WELL_2D, x, y)[0]
        rpos, ok = langevin.mc_placement(1234, 10 ** 10 + 7, 5000, en, beta, (-2.5, 2.5, -2.0, 2.0))
        assert ok.all()
        assert rel_err(pos[:2].T, rpos) < (1e-12 if prec == "fp64" else 1e-6)
        assert np.all(pos[2] == 0)
        e.close()
    # a box where nothing is acceptable reports every walker as failed (RuntimeError upstream)
    e = biasspec.make_ensemble(100, 2, 1, "fp32", potentials.DOUBLE_WELL_2D, None, dict(kind="none"))
    assert e.init_positions_mc((2.4, 2.5, 1.9, 2.0), 1 / (langevin.KB_PY * 300), ntrials=5) == 100
    e.close()


# ---------------------------------------------------------------------------------------------------
def test_eval_bias_golden_legendre1d(golden):
    for c in golden("legendre1d.json"):
        spec = dict(kind="leg1d", axis=c["axis"], degree=c["degree"], min=c["min"], max=c["max"], weights=c["weights"])
        for prec, tol, dt in PREC:
            e = biasspec.make_ensemble(1, 2, 0, prec, potentials.DOUBLE_WELL_2D, None, spec)
            pos = torch.tensor(c["positions"], dtype=dt, device=DEV)
            E, F = e.eval_bias(pos)
            # FP32: the degree-63 sum of O(1) terms carries ~K * eps absolute error
            scale = max(1.0, float(np.abs(c["weights"]).sum()))
            assert rel_err(E.cpu(), c["E"], scale) < tol, (c["tag"], prec)
            fscale = max(1.0, float(np.max(np.abs(c["F"]))))
            assert rel_err(F.cpu(), c["F"], fscale) < tol, (c["tag"], prec)
            # host-buffer path of the same entry point
            Eh, Fh = e.eval_bias(pos.cpu().numpy())
            assert np.array_equal(Eh, E.cpu().numpy()) and np.array_equal(Fh, F.cpu().numpy())
            e.close()


def test_eval_bias_golden_nn_pathcv_harmonic(golden):
    for c in golden("nnbasis1d.json"):
        spec = dict(kind="mlp", axis=c["axis"], min=c["min"], max=c["max"], sizes=c["sizes"],
                    act={"relu": 0, "tanh": 1}[c["act"]], theta=c["theta"])
        for prec, tol, dt in PREC:
            e = biasspec.make_ensemble(1, 2, 0, prec, potentials.SZABO_BEREZHKOVSKII, None, spec)
            E, F = e.eval_bias(torch.tensor(c["positions"], dtype=dt, device=DEV))
            es = max(1.0, float(np.max(np.abs(c["E"]))))
            fs = max(1.0, float(np.max(np.abs(c["F"]))))
            # FP32 can flip a ReLU that sits within rounding of its kink; allow isolated rows
            bad = np.abs(E.cpu().numpy() - c["E"]) / es > tol
            assert bad.sum() <= (0 if prec == "fp64" else 1), (c["tag"], prec)
            badf = np.abs(F.cpu().numpy() - c["F"]).max(1) / fs > tol
            assert badf.sum() <= (0 if prec == "fp64" else 2), (c["tag"], prec)
            e.close()
    for c in golden("pathcv.json"):
        spec = dict(kind="pathcv", degree=c["degree"], x_i=c["x_i"], y_i=c["y_i"], lam=c["lam"], weights=c["weights"], phi=c["phi"])
        for prec, tol, dt in PREC:
            e = biasspec.make_ensemble(1, 2, 0, prec, potentials.SLIP_BOND_2D, None, spec)
            E, F = e.eval_bias(torch.tensor(c["positions"], dtype=dt, device=DEV))
            # FP32 too: the exponents come from differences of squared distances taken as products (bias.cuh), so lam = 100
            # no longer amplifies the rounding of (x - x_i)^2 (round 1 needed 2e-3 here)
            assert rel_err(E.cpu(), c["E"], max(1.0, float(np.max(np.abs(c["E"]))))) < tol, (c["tag"], prec)
            assert rel_err(F.cpu(), c["F"], max(1.0, float(np.max(np.abs(c["F"]))))) < tol, (c["tag"], prec)
            e.close()
    c = golden("harmonic.json")
    e = biasspec.make_ensemble(1, 2, 0, "fp64", potentials.SLIP_BOND_2D, None, dict(kind="harmonic", axis=c["axis"], k=c["k"], s0=c["s0"]))
    E, F = e.eval_bias(torch.tensor(c["positions"], dtype=torch.float64, device=DEV))
    assert rel_err(E.cpu(), c["E"]) < 1e-12 and rel_err(F.cpu(), c["F"]) < 1e-12
    e.close()


def test_eval_bias_legendre2d_vs_oracle():
    """2D tensor-product basis against the oracle: 1e-6 (FP64) / 1e-4 (FP32) of the largest force for every degree up to
    31 x 31.  At degree 63 in FP32 the sum itself is ill conditioned -- |P'_63| reaches 2016 and the terms w_ij P'_i P_j
    cancel -- so each element is held to the forward-error bound of the evaluation instead,
        |dF_x| <= 4 (dx + dy + 2) eps_32 sum_ij |w_ij| |P'_i(x')| |P_j(y')| / half_width_x      (same for y, E),
    i.e. rounding error of a (dx + dy)-step recursion times the sum of the ABSOLUTE terms: the tightest statement an FP32
    plus the effect of the rounding of the scaled coordinate itself, 2 eps_32 (1 + |x'|) sum_ij |w_ij| |P''_i| |P_j| (and the
    cross term with P'_i P'_j for y'): the tightest statement an FP32 evaluation of this sum admits, element by element
    (round 1 allowed a flat 5e-4)."""
    from oracle import legendre as oleg
    rng = np.random.default_rng(5)
    pos = np.stack([rng.uniform(-3, 3, 4000), rng.uniform(-2.5, 2.5, 4000)], 1)
    pos[:4] = [[-2.5, 2.0], [2.5, -2.0], [2.6, 0.0], [0.0, -2.1]]      # clamp edges / outside
    eps32 = float(np.finfo(np.float32).eps)
    for dx, dy in [(7, 7), (15, 15), (31, 31), (63, 63), (3, 20), (20, 3), (0, 0), (1, 0)]:
        spec = rng_spec("leg2d", rng, deg_x=dx, deg_y=dy)
        Er, Fr = biasspec.oracle_energy_force(spec, pos)
        mm = spec["minmax"]
        sx, inx = oleg.scale_clamp(pos[:, 0], mm[0], mm[1])
        sy, iny = oleg.scale_clamp(pos[:, 1], mm[2], mm[3])
        Px, dPx = oleg.legendre_table(sx, dx)
        Py, dPy = oleg.legendre_table(sy, dy)
        aw = np.abs(np.asarray(spec["weights"]))
        L = np.polynomial.legendre
        d2Px = np.abs(L.legval(sx, L.legder(np.eye(dx + 1), 2, axis=0))) if dx >= 2 else np.zeros_like(Px)
        d2Py = np.abs(L.legval(sy, L.legder(np.eye(dy + 1), 2, axis=0))) if dy >= 2 else np.zeros_like(Py)
        if dx >= 2:
            d2Px = np.concatenate([d2Px, np.zeros((dx + 1 - d2Px.shape[0], len(sx)))], 0) if d2Px.shape[0] < dx + 1 else d2Px
        if dy >= 2:
            d2Py = np.concatenate([d2Py, np.zeros((dy + 1 - d2Py.shape[0], len(sy)))], 0) if d2Py.shape[0] < dy + 1 else d2Py
        ein = lambda A, B: np.einsum("ij,in,jn->n", aw, np.abs(A), np.abs(B))
        hx, hy = (mm[1] - mm[0]) / 2, (mm[3] - mm[2]) / 2
        dsx, dsy = 2 * eps32 * (1 + np.abs(sx)), 2 * eps32 * (1 + np.abs(sy))
        condE = ein(Px, Py)
        condF = np.stack([ein(dPx, Py) * inx / hx, ein(Px, dPy) * iny / hy], 1)
        inE = ein(dPx, Py) * dsx + ein(Px, dPy) * dsy
        inF = np.stack([(ein(d2Px, Py) * dsx + ein(dPx, dPy) * dsy) * inx / hx, (ein(dPx, dPy) * dsx + ein(Px, d2Py) * dsy) * iny / hy], 1)
        for prec, tol, dt in PREC:
            e = biasspec.make_ensemble(1, 2, 0, prec, potentials.DOUBLE_WELL_2D, None, spec)
            E, F = e.eval_bias(torch.tensor(pos, dtype=dt, device=DEV))
            E, F = E.cpu().numpy().astype(np.float64), F.cpu().numpy().astype(np.float64)
            scale = max(1.0, float(np.abs(spec["weights"]).sum()))
            fscale = max(1.0, float(np.max(np.abs(Fr))))
            if max(dx, dy) < 40 or prec == "fp64":
                assert rel_err(E, Er, scale) < tol, (dx, dy, prec)
                assert rel_err(F, Fr, fscale) < tol, (dx, dy, prec)
            else:
                g = 4 * (dx + dy + 2) * eps32
                assert np.all(np.abs(E - Er) <= g * condE + inE + 1e-30), (dx, dy, prec, float(np.max(np.abs(E - Er) / (g * condE + inE))))
                assert np.all(np.abs(F - Fr) <= g * condF + inF + 1e-30), (dx, dy, prec, float(np.max(np.abs(F - Fr) / (g * condF + inF + 1e-30))))
                assert rel_err(F, Fr, fscale) < 5e-4                   # what that amounts to on this data set
            e.close()


# ---------------------------------------------------------------------------------------------------
TRAJ_CASES = [
    # (tag, potential, params, bias builder, start box, variants)
    ("dw_none", potentials.DOUBLE_WELL_2D, None, lambda r: dict(kind="none"), (-1.6, -1.2, -0.3, 0.3), [-1]),
    ("dw_leg1d5", potentials.DOUBLE_WELL_2D, None, lambda r: rng_spec("leg1d", r, degree=5), (-1.6, -1.2, -0.3, 0.3), [-1]),
    # variants: -1 default table (FP32: tensor-core contraction except the static 8 x 8 double-well kernel), 7 tensor cores forced, 6 CUDA cores forced
    ("dw_leg2d8", potentials.DOUBLE_WELL_2D, None, lambda r: rng_spec("leg2d", r, deg_x=7, deg_y=7), (-1.6, -1.2, -0.3, 0.3), [-1, 7]),
    ("dw_leg2d16", potentials.DOUBLE_WELL_2D, None, lambda r: rng_spec("leg2d", r, deg_x=15, deg_y=15), (-1.6, -1.2, -0.3, 0.3), [-1, 1, 99, 6]),
    ("dw_leg2d32", potentials.DOUBLE_WELL_2D, None, lambda r: rng_spec("leg2d", r, deg_x=31, deg_y=31, scale=1.0), (-1.6, -1.2, -0.3, 0.3), [-1, 6]),
    ("dw_leg2d64", potentials.DOUBLE_WELL_2D, None, lambda r: rng_spec("leg2d", r, deg_x=63, deg_y=63, scale=0.5), (-1.6, -1.2, -0.3, 0.3), [-1, 6]),
    ("dw_leg2d_5x11", potentials.DOUBLE_WELL_2D, None, lambda r: rng_spec("leg2d", r, deg_x=4, deg_y=10), (-1.6, -1.2, -0.3, 0.3), [-1, 7]),
    ("dw_leg2d_17x3", potentials.DOUBLE_WELL_2D, None, lambda r: rng_spec("leg2d", r, deg_x=16, deg_y=2), (-1.6, -1.2, -0.3, 0.3), [-1, 6]),
    ("dw_leg2d_2x50", potentials.DOUBLE_WELL_2D, None, lambda r: rng_spec("leg2d", r, deg_x=1, deg_y=49, scale=1.0), (-1.6, -1.2, -0.3, 0.3), [-1]),
    ("slip_none", potentials.SLIP_BOND_2D, None, lambda r: dict(kind="none"), (-3.7, -3.2, -3.7, -3.2), [-1]),
    ("slip_leg2d16", potentials.SLIP_BOND_2D, None, lambda r: rng_spec("leg2d", r, deg_x=15, deg_y=15, minmax=(-6.0, 6.0, -6.0, 6.0), scale=1.0), (-3.7, -3.2, -3.7, -3.2), [-1, 2]),
    ("mb_leg2d_40x24", potentials.MULLER_BROWN, None, lambda r: rng_spec("leg2d", r, deg_x=39, deg_y=23, minmax=(-1.5, 1.0, 0.0, 2.0), scale=0.3), (-0.6, -0.5, 1.4, 1.5), [-1]),
    ("sb_leg2d_20x36", potentials.SZABO_BEREZHKOVSKII, None, lambda r: rng_spec("leg2d", r, deg_x=19, deg_y=35, minmax=(-4.0, 4.0, -4.0, 4.0), scale=0.5), (2.0, 2.4, 2.0, 2.4), [-1]),
    ("slip_leg1d12", potentials.SLIP_BOND_2D, None, lambda r: rng_spec("leg1d", r, degree=12, min=-3.0, max=6.0, scale=1.0), (-3.7, -3.2, -3.7, -3.2), [-1]),
    ("slip_forced_leg1d7y", potentials.SLIP_BOND_2D, (1.5, -0.5, 0.8, 4.0, 3.0, 0.2, 2.5), lambda r: rng_spec("leg1d", r, degree=7, axis='y', min=-4.0, max=6.0, scale=1.0), (-3.0, -2.5, -3.0, -2.5), [-1]),
    ("sb_none", potentials.SZABO_BEREZHKOVSKII, None, lambda r: dict(kind="none"), (2.0, 2.4, 2.0, 2.4), [-1]),
    # ReLU nets, variants: -1 = FP32 default, the piecewise-constant dV/ds table (BiasPwl1D; FP64 handles: the network itself), 8 = the table
    # forced (FP64 too), 7 = the H x H layer on the tensor cores, 6 = CUDA cores
    ("sb_mlp48x3", potentials.SZABO_BEREZHKOVSKII, None, lambda r: rng_spec("mlp", r, sizes=[48, 48, 48]), (2.0, 2.4, 2.0, 2.4), [-1, 6, 7, 8]),
    # ReLU nets with one H x H layer run the mask-operand tensor-core kernel (H padded to 32 / 48 / 64)
    ("sb_mlp_relu_32", potentials.SZABO_BEREZHKOVSKII, None, lambda r: rng_spec("mlp", r, sizes=[30, 32, 28]), (2.0, 2.4, 2.0, 2.4), [-1, 6, 7, 8]),
    ("catch_mlp_relu_64", potentials.CATCH_BOND_2D, None, lambda r: rng_spec("mlp", r, sizes=[64, 60, 64], scale=0.25), (-3.7, -3.2, -3.7, -3.2), [-1, 6, 7, 8]),
    ("sb_mlp128x2", potentials.SZABO_BEREZHKOVSKII, None, lambda r: rng_spec("mlp", r, sizes=[128, 128]), (2.0, 2.4, 2.0, 2.4), [-1, 6, 8]),
    ("sb_mlp16", potentials.SZABO_BEREZHKOVSKII, None, lambda r: rng_spec("mlp", r, sizes=[16]), (2.0, 2.4, 2.0, 2.4), [-1]),
    ("catch_mlp_tanh", potentials.CATCH_BOND_2D, None, lambda r: rng_spec("mlp", r, sizes=[20, 30, 24], act=1), (-3.7, -3.2, -3.7, -3.2), [-1, 6]),
    ("sb_mlp_64_tanh", potentials.SZABO_BEREZHKOVSKII, None, lambda r: rng_spec("mlp", r, sizes=[64, 64, 56], act=1, scale=0.25), (2.0, 2.4, 2.0, 2.4), [-1, 6]),
    ("catch_mlp_deep", potentials.CATCH_BOND_2D, None, lambda r: rng_spec("mlp", r, sizes=[8, 12, 10, 6]), (-3.7, -3.2, -3.7, -3.2), [-1]),
    ("mb_leg1d", potentials.MULLER_BROWN, None, lambda r: rng_spec("leg1d", r, degree=9, min=-1.5, max=1.0, scale=0.5), (-0.6, -0.5, 1.4, 1.5), [-1]),
    ("sw1d_leg1d", potentials.SINGLE_WELL_1D, None, lambda r: rng_spec("leg1d", r, degree=3, min=-2.0, max=2.0), (-0.3, 0.3, -0.01, 0.01), [-1]),
    ("dw1d_none", potentials.DOUBLE_WELL_1D, None, lambda r: dict(kind="none"), (-1.5, -1.2, -0.01, 0.01), [-1]),
    ("sw2d_pathcv", potentials.SINGLE_WELL_2D, None, lambda r: dict(kind="pathcv", degree=4, x_i=np.linspace(-2, 2, 30), y_i=np.linspace(-1, 1.5, 30), lam=20.0, weights=r.standard_normal(5), phi=3.0), (-0.5, 0.5, -0.5, 0.5), [-1]),
    ("sw2d_harmonic", potentials.SINGLE_WELL_2D, None, lambda r: dict(kind="harmonic", axis='y', k=12.5, s0=0.4), (-0.5, 0.5, -0.5, 0.5), [-1]),
]


@pytest.mark.parametrize("case", TRAJ_CASES, ids=[c[0] for c in TRAJ_CASES])
def test_trajectory_parity_injected_noise(case):
    """1e-6 relative over 10^3 steps in FP64 (1e-4 in FP32 over 200 steps inside a basin)."""
    tag, pk, pp, mk, box, variants = case
    rng = np.random.default_rng(zlib.crc32(tag.encode()))
    spec = mk(rng)
    n = 257                                  # not a multiple of the block size: exercises the tail
    k = langevin.constants()
    x0 = np.stack([rng.uniform(box[0], box[1], n), rng.uniform(box[2], box[3], n)], 1)
    v0 = rng.standard_normal((n, 2)) * k["sigma_v"]
    f = biasspec.total_force_fn(pk, pp, spec)
    for prec, tol, dt in PREC:
        nsteps = 1000 if prec == "fp64" else 200
        noise = rng.standard_normal((nsteps, 2, n))
        xr, vr = langevin.run_injected(x0, v0, f, k, noise)
        for variant in variants:
            e = biasspec.make_ensemble(n, 2, 0, prec, pk, pp, spec)
            e.set_tuning(variant)
            e.set_state(torch.tensor(x0.T.copy(), dtype=dt, device=DEV), torch.tensor(v0.T.copy(), dtype=dt, device=DEV))
            tn = torch.tensor(noise, dtype=dt, device=DEV)
            # odd chunking exercises the odd-head / even-tail handling of the two-step loop
            cuts = [0, 333, 334, nsteps] if prec == "fp64" else [0, 77, nsteps]
            for a, b in zip(cuts[:-1], cuts[1:]):
                e.step(b - a, tn[a:b].contiguous())
            assert e.step_counter == nsteps
            x, v = e.get_state(host=True)
            assert rel_err(x.T, xr) < tol, (tag, prec, variant, rel_err(x.T, xr))
            assert rel_err(v.T, vr) < tol * (1 if prec == "fp64" else 10), (tag, prec, variant)
            e.close()


def test_tensor_core_step_force_vs_oracle():
    """The tensor-core contraction of the step kernel (csrc/langevin_tc.cu: FP16 hi + lo operand split, three MMAs per product,
    in-kernel power-of-two weight scale) against the FP64 oracle, element by element: one noiseless step from rest gives
    v = -b grad(U + V), for basis shapes that exercise every padded tile size (16 / 32 / 64 per axis), weights over ten
    decades of magnitude, walkers inside, on the edges of and outside the box.  1e-4 of the largest force is the north-star's
    FP32 bound; measured 2e-7."""
    rng = np.random.default_rng(77)
    k = langevin.constants()
    n = 3001
    pos = np.stack([rng.uniform(-2.7, 2.7, n), rng.uniform(-2.2, 2.2, n)], 1)
    pos[:4] = [[-2.5, 2.0], [2.5, -2.0], [2.6, 0.0], [0.0, -2.1]]
    zero = torch.zeros((1, 2, n), dtype=torch.float32, device=DEV)
    for (dx, dy), variant in [((63, 63), -1), ((32, 32), -1), ((1, 49), -1), ((16, 2), -1), ((39, 23), -1), ((15, 15), 7), ((7, 7), 7), ((4, 10), 7)]:
        for scale in (1e-6, 1.0, 1e4):
            spec = rng_spec("leg2d", rng, deg_x=dx, deg_y=dy, scale=scale)
            f = biasspec.total_force_fn(potentials.DOUBLE_WELL_2D, None, spec)
            F = f(pos)
            e = biasspec.make_ensemble(n, 2, 0, "fp32", potentials.DOUBLE_WELL_2D, None, spec)
            e.set_tuning(variant)
            e.set_state(torch.tensor(pos.T.copy(), dtype=torch.float32, device=DEV), torch.zeros((2, n), dtype=torch.float32, device=DEV))
            e.step(1, zero)
            _, v = e.get_state(host=True)
            got = v.T.astype(np.float64) / k["b_over_m"]
            tol = 1e-4 if max(dx, dy) < 40 else 5e-4          # degree >= 40: the FP32 forward-error bound of the sum itself (see above)
            assert rel_err(got, F, max(1.0, float(np.max(np.abs(F))))) < tol, (dx, dy, variant, scale)
            # the same launch on the CUDA cores agrees with it at FP32 rounding level
            e.set_tuning(6)
            e.set_state(torch.tensor(pos.T.copy(), dtype=torch.float32, device=DEV), torch.zeros((2, n), dtype=torch.float32, device=DEV))
            e.step(1, zero)
            _, v2 = e.get_state(host=True)
            assert rel_err(v2.T.astype(np.float64) / k["b_over_m"], F, max(1.0, float(np.max(np.abs(F))))) < tol
            e.close()


@pytest.mark.parametrize("shape", [(15, 15), (31, 31), (63, 63), (16, 2), "mlp48"], ids=["16x16", "32x32", "64x64", "17x3_two_sets_per_thread", "relu_mlp48"])
def test_tensor_core_step_every_walker_once(shape):
    """How the tensor-core launch deals the sets of 128 walkers to its CTAs (whole rounds of full CTAs, the remainder in small CTAs
    spread over the SMs, the MLP kernel likewise): for ensemble sizes around one and two rounds of the grid, with ragged ends, a few
    RNG-driven steps (Philox keyed on the global walker id) must move EVERY walker exactly as the CUDA-core kernel does -- a set
    skipped, done twice or mapped to the wrong ids would show as walkers that did not move or moved with another walker's noise."""
    rng = np.random.default_rng(5)
    if shape == "mlp48":
        spec = rng_spec("mlp", rng, sizes=[48, 48, 48])
    else:
        spec = rng_spec("leg2d", rng, deg_x=shape[0], deg_y=shape[1], scale=1.0)
    n_sm = torch.cuda.get_device_properties(0).multi_processor_count
    probe = biasspec.make_ensemble(128, 2, 0, "fp32", potentials.DOUBLE_WELL_2D, None, spec)
    if shape == "mlp48":
        probe.set_tuning(7)
    wave = probe.wave_walkers()
    probe.close()
    assert wave > 0 and wave % (128 * n_sm) == 0
    for n in (1, 127, 129, wave - 128 - 1, wave, wave + 1, wave + 300, 2 * wave + 128 * n_sm + 77):
        x0 = np.stack([rng.uniform(-1.6, 1.6, n), rng.uniform(-1.0, 1.0, n)], 1)
        v0 = np.zeros((n, 2))
        out = []
        for variant in ((7, 6) if shape == "mlp48" else (-1, 6)):
            e = biasspec.make_ensemble(n, 2, 123, "fp32", potentials.DOUBLE_WELL_2D, None, spec)
            e.set_walker_offset(10 ** 9)
            e.set_tuning(variant)
            e.set_state(torch.tensor(x0.T.copy(), dtype=torch.float32, device=DEV), torch.tensor(v0.T.copy(), dtype=torch.float32, device=DEV))
            e.step(3)
            e.step(2)
            out.append(e.get_state(host=True))
            e.close()
        (xa, va), (xb, vb) = out
        assert np.all(np.abs(xa.T - x0) > 0), (shape, n)                # every walker moved
        assert np.max(np.abs(xa - xb)) < 2e-5 and np.max(np.abs(va - vb)) < 2e-3, (shape, n, np.max(np.abs(xa - xb)))


@pytest.mark.parametrize("deg", [7, 15, 31, 63], ids=["8x8_cuda_cores", "16x16_tensor_cores", "32x32_tensor_cores", "64x64_tensor_cores"])
def test_trajectory_parity_philox_and_chunking(deg):
    """RNG-driven run against the oracle's Philox stream; launches of any length give identical bits."""
    rng = np.random.default_rng(11)
    spec = rng_spec("leg2d", rng, deg_x=deg, deg_y=deg, scale=3.0 if deg <= 15 else 1.0)
    k = langevin.constants()
    n, gid0, seed = 300, 2 ** 33 + 5, 20221019
    x0 = np.stack([rng.uniform(-1.6, -1.2, n), rng.uniform(-0.3, 0.3, n)], 1)
    v0 = rng.standard_normal((n, 2)) * k["sigma_v"]
    f = biasspec.total_force_fn(potentials.DOUBLE_WELL_2D, None, spec)
    xr, vr = langevin.run_philox(x0, v0, f, k, seed, gid0, 0, 301)
    for prec, tol, dt in PREC:
        outs = []
        for cuts in ([0, 301], [0, 1, 2, 150, 151, 301], [0, 7, 300, 301]):
            e = biasspec.make_ensemble(n, 2, seed, prec, potentials.DOUBLE_WELL_2D, None, spec)
            e.set_walker_offset(gid0)
            e.set_state(torch.tensor(x0.T.copy(), dtype=dt, device=DEV), torch.tensor(v0.T.copy(), dtype=dt, device=DEV))
            for a, b in zip(cuts[:-1], cuts[1:]):
                e.step(b - a)
            outs.append(e.get_state(host=True))
            e.close()
        assert rel_err(outs[0][0].T, xr) < (1e-9 if prec == "fp64" else 2e-3)
        for o in outs[1:]:
            assert np.array_equal(o[0], outs[0][0]) and np.array_equal(o[1], outs[0][1])


def test_full_size_sharding_invariance():
    """BASELINE.json configs[1] at full size (10^6 walkers, 16 x 16 bias): the ensemble cut into two shards with
    global walker offsets reproduces the single-handle run bit for bit (Philox counters use global ids), and the
    shard moments add up to the ensemble's (FP64 fixed-order reductions): the properties the multi-GPU path
    relies on.  Also linearity of the bias in its coefficients at this size."""
    rng = np.random.default_rng(21)
    n, half, seed, nsteps = 1000000, 400000, 20221019, 24
    spec = rng_spec("leg2d", rng, deg_x=15, deg_y=15)
    spec2 = dict(spec, weights=rng.standard_normal(spec["weights"].shape) * 0.3)
    both = dict(spec, weights=spec["weights"] + spec2["weights"])

    def run(count, offset, sp):
        e = biasspec.make_ensemble(count, 2, seed, "fp32", potentials.DOUBLE_WELL_2D, None, sp)
        e.set_walker_offset(offset)
        assert e.init_positions_mc((-2.5, 2.5, -2.0, 2.0), 1.0 / (8.3145e-3 * 300)) == 0
        e.init_velocities()
        e.step(nsteps)
        x, v = e.get_state()
        sw, sf, _ = e.moments()
        return e, x, v, sw.cpu().numpy(), sf.cpu().numpy()

    e, x, v, sw, sf = run(n, 0, spec)
    ea, xa, va, swa, sfa = run(half, 0, spec)
    eb, xb, vb, swb, sfb = run(n - half, half, spec)
    assert torch.equal(x[:, :half], xa) and torch.equal(x[:, half:], xb)
    assert torch.equal(v[:, :half], va) and torch.equal(v[:, half:], vb)
    assert sw[0] == n and swa[0] + swb[0] == n and abs(sf[0] - n) < 1e-6           # f_0 = P_0 P_0 = 1
    assert np.max(np.abs(sfa + sfb - sf)) < 1e-6 * n
    # linearity: E(w1 + w2) = E(w1) + E(w2), F likewise (FP32 evaluation, 10^6 rows)
    pos = x.T.contiguous()
    E1, F1 = e.eval_bias(pos)
    biasspec.apply(e, spec2)
    E2, F2 = e.eval_bias(pos)
    biasspec.apply(e, both)
    E12, F12 = e.eval_bias(pos)
    scale = float(torch.max(torch.abs(E1)) + torch.max(torch.abs(E2)))
    assert float(torch.max(torch.abs(E12 - (E1 + E2)))) < 2e-5 * scale
    fscale = float(torch.max(torch.abs(F1)) + torch.max(torch.abs(F2)))
    assert float(torch.max(torch.abs(F12 - (F1 + F2)))) < 2e-5 * fscale
    for h in (e, ea, eb):
        h.close()


def test_empty_and_ragged_inputs():
    """Empty row sets (eval_bias, moments, histogram), zero-length launches and ensembles that do not fill a
    thread block / a pair of walkers / a tile: n = 1, 2, 31, 129, 257 against the oracle."""
    rng = np.random.default_rng(5)
    k = langevin.constants()
    specs = [rng_spec("leg1d", rng, degree=5), rng_spec("leg2d", rng, deg_x=15, deg_y=15), rng_spec("leg2d", rng, deg_x=4, deg_y=6)]
    for spec in specs:
        f = biasspec.total_force_fn(potentials.DOUBLE_WELL_2D, None, spec)
        for n in (1, 2, 31, 129, 257):
            x0 = np.stack([rng.uniform(-1.6, -1.2, n), rng.uniform(-0.3, 0.3, n)], 1)
            v0 = rng.standard_normal((n, 2)) * k["sigma_v"]
            xr, vr = langevin.run_philox(x0, v0, f, k, 3, 0, 0, 9)
            e = biasspec.make_ensemble(n, 2, 3, "fp64", potentials.DOUBLE_WELL_2D, None, spec)
            e.set_state(torch.tensor(x0.T.copy(), dtype=torch.float64, device=DEV), torch.tensor(v0.T.copy(), dtype=torch.float64, device=DEV))
            e.step(0)                                   # zero-length launch: nothing happens
            e.step(9)
            x, v = e.get_state(host=True)
            assert rel_err(x.T, xr) < 1e-9 and rel_err(v.T, vr) < 1e-9, (spec["kind"], n)
            sw, sf, sff = e.moments(cov=True)
            fr = biasspec.oracle_basis(spec, x.T)
            assert float(sw.cpu()[0]) == n and np.max(np.abs(sf.cpu().numpy() - fr.sum(0))) < 1e-9 * max(1, n)
            assert np.max(np.abs(sff.cpu().numpy() - fr.T @ fr)) < 1e-9 * max(1, n)
            e.close()
        # empty caller arrays
        e = biasspec.make_ensemble(4, 2, 3, "fp32", potentials.DOUBLE_WELL_2D, None, spec)
        empty = torch.empty((0, 2), dtype=torch.float32, device=DEV)
        E, F = e.eval_bias(empty)
        assert E.shape == (0,) and F.shape == (0, 2)
        sw, sf, sff = e.moments(pos=empty, cov=True)
        assert float(sw.cpu()[0]) == 0 and not sf.cpu().numpy().any() and not sff.cpu().numpy().any()
        assert float(e.histogram((-1.0, 1.0), 8, axis=0, pos=empty).sum().cpu()) == 0
        e.close()


def test_int64_indexing_beyond_2_31_walkers():
    """Maximum-size case: 2^31 + 4096 walkers (34 GB of FP32 state on the 180 GB B200).  Placement, the packed step
    kernel, the one-pass moments and the histogram index with 64 bits: every walker is counted, and the last 4096
    walkers equal a 4096-walker handle with the same global ids bit for bit."""
    free, _ = torch.cuda.mem_get_info()
    n = 2 ** 31 + 4096
    if free < 90e9:
        pytest.skip("needs ~75 GB of free device memory")
    kT = 8.3145e-3 * 300
    w = np.array([0.3, -1.2, 0.8, 0.1, -0.4, 0.2])

    def make(count, offset):
        e = _lib.Ensemble(count, dims=2, seed=99, precision="fp32")
        e.set_integrator(1.0, 8.31446261815324e-3 * 300, 20.0, 0.01)
        e.set_potential(_lib.POT_DOUBLE_WELL_2D)
        e.set_walker_offset(offset)
        e.set_bias_legendre1d(0, 5, -2.5, 2.5, w)
        assert e.init_positions_mc((-2.5, 2.5, -2.0, 2.0), 1.0 / kT) == 0
        e.init_velocities()
        e.step(6)
        return e

    big = make(n, 0)
    sw, sf, sff = big.moments(cov=True)
    assert float(sw.cpu()[0]) == float(n) and float(sff.cpu()[0, 0]) == float(n)
    counts = big.histogram((-10.0, 10.0), 64, axis=0)
    assert float(counts.sum().cpu()) == float(n)
    x, v = big.get_state()
    tail = make(4096, n - 4096)
    xt, vt = tail.get_state()
    assert torch.equal(x[:, n - 4096:], xt) and torch.equal(v[:, n - 4096:], vt)
    assert bool(torch.isfinite(x[:, -1]).all()) and float(x[0, n - 1]) != 0.0
    del x, v
    big.close()
    tail.close()
    torch.cuda.empty_cache()


def test_general_kernel_3d_middle_record():
    """dims = 3 (z simulated, 1000 z^2), both schemes, trajectory recording (frame t before step t)."""
    rng = np.random.default_rng(3)
    spec = rng_spec("leg1d", rng, degree=5)
    k = langevin.constants()
    n, nsteps = 70, 300
    x0 = np.concatenate([np.stack([rng.uniform(-1.6, -1.2, n), rng.uniform(-0.3, 0.3, n)], 1), rng.standard_normal((n, 1)) * 0.02], 1)
    v0 = rng.standard_normal((n, 3)) * k["sigma_v"]
    f = biasspec.total_force_fn(potentials.DOUBLE_WELL_2D, None, spec)
    noise = rng.standard_normal((nsteps, 3, n))
    for scheme, sid in (("openmm_langevin", langevin.SCHEME_OPENMM), ("middle", langevin.SCHEME_MIDDLE)):
        xr, vr, frames = langevin.run_injected(x0, v0, f, k, noise, scheme=sid, record=True)
        e = biasspec.make_ensemble(n, 3, 0, "fp64", potentials.DOUBLE_WELL_2D, None, spec, scheme=scheme)
        e.set_state(torch.tensor(x0.T.copy(), device=DEV), torch.tensor(v0.T.copy(), device=DEV))
        traj = e.step_record(nsteps, every=7, n_rec=5, noise=torch.tensor(noise, device=DEV))
        x, v = e.get_state(host=True)
        assert rel_err(x.T, xr) < 1e-6 and rel_err(v.T, vr) < 1e-6, scheme
        want = frames[::7, :5, :].transpose(0, 2, 1)
        assert traj.shape == (43, 3, 5)
        assert rel_err(traj.cpu(), want) < 1e-6, scheme
        e.close()
    # Philox-driven 3D run
    xr, vr = langevin.run_philox(x0, v0, f, k, 99, 0, 0, 50)
    e = biasspec.make_ensemble(n, 3, 99, "fp64", potentials.DOUBLE_WELL_2D, None, spec)
    e.set_state(torch.tensor(x0.T.copy(), device=DEV), torch.tensor(v0.T.copy(), device=DEV))
    e.step(50)
    x, v = e.get_state(host=True)
    assert rel_err(x.T, xr) < 1e-9
    e.close()


def test_fast_and_general_kernels_agree():
    rng = np.random.default_rng(4)
    spec = rng_spec("leg2d", rng, deg_x=15, deg_y=15)
    n = 500
    outs = []
    for variant in (-1, 99):
        e = biasspec.make_ensemble(n, 2, 77, "fp64", potentials.DOUBLE_WELL_2D, None, spec)
        e.set_tuning(variant)
        e.init_positions_mc((-2.5, 2.5, -2, 2), 0.4)
        e.init_velocities()
        e.step(101)
        outs.append(e.get_state(host=True))
        e.close()
    assert rel_err(outs[0][0], outs[1][0]) < 1e-9


# ---------------------------------------------------------------------------------------------------
def test_energies_vs_oracle():
    rng = np.random.default_rng(8)
    spec = rng_spec("leg1d", rng, degree=5)
    k = langevin.constants()
    n = 1000
    for dims in (2, 3):
        x0 = np.stack([rng.uniform(-2, 2, n), rng.uniform(-1, 1, n), rng.standard_normal(n) * 0.03][:dims], 1)
        v0 = rng.standard_normal((n, dims)) * k["sigma_v"]
        f = biasspec.total_force_fn(potentials.DOUBLE_WELL_2D, None, spec)
        PEr = potentials.energy(potentials.DOUBLE_WELL_2D, x0) + biasspec.oracle_energy_force(spec, x0)[0]
        KEr = langevin.kinetic_energy(v0, f(x0), k)
        for prec, tol, dt in PREC:
            e = biasspec.make_ensemble(n, dims, 0, prec, potentials.DOUBLE_WELL_2D, None, spec)
            e.set_state(torch.tensor(x0.T.copy(), dtype=dt, device=DEV), torch.tensor(v0.T.copy(), dtype=dt, device=DEV))
            PE, KE = e.energies()
            assert rel_err(PE.cpu(), PEr, 10.0) < tol and rel_err(KE.cpu(), KEr) < tol
            PEh, KEh = e.energies(host=True)
            assert np.array_equal(PEh, PE.cpu().numpy())
            e.close()


MOMENT_CASES = [
    ("leg1d5", lambda r: rng_spec("leg1d", r, degree=5), True),
    ("leg1d12y", lambda r: rng_spec("leg1d", r, degree=12, axis='y', min=-2.0, max=2.0), True),
    ("leg1d63", lambda r: rng_spec("leg1d", r, degree=63), True),
    ("leg2d16", lambda r: rng_spec("leg2d", r, deg_x=15, deg_y=15), True),
    ("leg2d_5x9", lambda r: rng_spec("leg2d", r, deg_x=4, deg_y=8), True),
    ("leg2d32", lambda r: rng_spec("leg2d", r, deg_x=31, deg_y=31), False),
    # K > 256: register-tiled first moments (4 x 4 functions per thread); ragged edges and the largest basis
    ("leg2d_19x37", lambda r: rng_spec("leg2d", r, deg_x=18, deg_y=36), False),
    ("leg2d64", lambda r: rng_spec("leg2d", r, deg_x=63, deg_y=63), False),
    ("pathcv", lambda r: dict(kind="pathcv", degree=6, x_i=np.linspace(-2, 2, 30), y_i=np.linspace(-1, 1.5, 30), lam=20.0, weights=r.standard_normal(7), phi=0.0), True),
    ("mlp48x3", lambda r: rng_spec("mlp", r, sizes=[48, 48, 48]), False),
    ("mlp128x2", lambda r: rng_spec("mlp", r, sizes=[128, 128]), False),          # the reference example's net (deepves...:102)
    ("mlp_32_64_tanh", lambda r: rng_spec("mlp", r, sizes=[32, 64], act=1), False),
    ("mlp_tanh", lambda r: rng_spec("mlp", r, sizes=[7, 9], act=1), False),
    ("mlp_one", lambda r: rng_spec("mlp", r, sizes=[16]), False),
]


@pytest.mark.parametrize("case", MOMENT_CASES, ids=[c[0] for c in MOMENT_CASES])
def test_moments_vs_oracle(case):
    tag, mk, cov = case
    rng = np.random.default_rng(zlib.crc32(tag.encode()))
    spec = mk(rng)
    n = 3001
    pos = np.stack([rng.uniform(-3, 3, n), rng.uniform(-2.5, 2.5, n)], 1)
    w = rng.uniform(0, 2, n)
    fr = biasspec.oracle_basis(spec, pos)
    from oracle import ves
    for prec, tol, dt in PREC:
        e = biasspec.make_ensemble(n, 2, 0, prec, potentials.DOUBLE_WELL_2D, None, spec)
        assert e.basis_size == fr.shape[1]
        tpos = torch.tensor(pos, dtype=dt, device=DEV)
        tw = torch.tensor(w, dtype=dt, device=DEV)
        e.set_state(tpos.T.contiguous(), None)
        for weights, tws in ((None, None), (w, tw)):
            sw_r, sf_r, sff_r = ves.moments(fr, weights)
            # (1) the handle's own walkers, (2) caller rows on the device, (3) caller rows on the host
            results = [e.moments(row_weights=tws, cov=cov), e.moments(pos=tpos, row_weights=tws, cov=cov),
                       e.moments(pos=pos, row_weights=weights, cov=cov)]
            for sw, sf, sff in results:
                sw, sf = np.asarray(sw.cpu() if hasattr(sw, "cpu") else sw), np.asarray(sf.cpu() if hasattr(sf, "cpu") else sf)
                assert abs(sw[0] - sw_r) / sw_r < (1e-12 if prec == "fp64" else 1e-6)
                scale = np.maximum(np.abs(fr).T @ (np.ones(n) if weights is None else weights), 1.0)
                assert np.max(np.abs(sf - sf_r) / scale) < tol, (tag, prec)
                if cov:
                    sff = np.asarray(sff.cpu() if hasattr(sff, "cpu") else sff)
                    assert np.array_equal(sff, sff.T)
                    assert np.max(np.abs(sff - sff_r)) / max(1.0, np.max(np.abs(sff_r))) < tol, (tag, prec)
        e.close()


@pytest.mark.parametrize("case", [c for c in MOMENT_CASES if c[0] in ("leg1d12y", "leg2d16", "leg2d_5x9", "leg2d32", "leg2d_19x37", "leg2d64")],
                         ids=lambda c: c[0])
def test_moments_inside_domain(case):
    """pyves_set_moment_domain(1): rows outside the basis interval get weight 0 (conditional averages on
    the support of the target), on the CUDA-core kernels and on the tcgen05 covariance (leg2d16)."""
    tag, mk, cov = case
    rng = np.random.default_rng(zlib.crc32(tag.encode()) + 1)
    spec = mk(rng)
    n = 4099
    pos = np.stack([rng.uniform(-3.2, 3.2, n), rng.uniform(-2.6, 2.6, n)], 1)
    w = rng.uniform(0, 2, n)
    fr = biasspec.oracle_basis(spec, pos)
    from oracle import ves
    for prec, tol, dt in PREC:
        e = biasspec.make_ensemble(n, 2, 0, prec, potentials.DOUBLE_WELL_2D, None, spec)
        p_dev = pos.astype(e.np_dtype).astype(np.float64)              # mask what the device sees
        inside = biasspec.oracle_inside(spec, p_dev)
        assert 0.2 * n < inside.sum() < 0.9 * n
        tpos = torch.tensor(pos, dtype=dt, device=DEV)
        e.set_state(tpos.T.contiguous(), None)
        e.set_moment_domain(True)
        for weights in (None, w):
            tws = None if weights is None else torch.tensor(weights, dtype=dt, device=DEV)
            wm = inside.astype(np.float64) * (1.0 if weights is None else weights)
            sw_r, sf_r, sff_r = ves.moments(fr, wm)
            sw, sf, sff = e.moments(row_weights=tws, cov=cov)
            assert abs(float(sw.cpu()[0]) - sw_r) / sw_r < (1e-12 if prec == "fp64" else 1e-6)
            scale = np.maximum(np.abs(fr).T @ wm, 1.0)
            assert np.max(np.abs(sf.cpu().numpy() - sf_r) / scale) < tol, (tag, prec)
            if cov:
                assert np.max(np.abs(sff.cpu().numpy() - sff_r)) / max(1.0, np.max(np.abs(sff_r))) < max(tol, 2e-3 if prec == "fp32" else 0), (tag, prec)
        e.set_moment_domain(False)
        sw, _, _ = e.moments()
        assert float(sw.cpu()[0]) == n
        e.close()


@pytest.mark.parametrize("sizes", [[48, 48, 48], [128, 128], [30, 32, 28]], ids=["48x3", "128x2", "30_32_28"])
def test_moments_relu_mlp_through_the_pieces(sizes):
    """Parameter-gradient moments of a ReLU NNBasis1D through the pieces of the piecewise-linear network (pyves_set_tuning 8; the FP32
    default for large ensembles): dV/dtheta is affine in the input on every piece, so binning the rows and running the gradient
    kernel on one pseudo row per piece gives the same sums as the per-row pass -- against the FP64 oracle, with and without row
    weights, with the in-box restriction, FP32 and FP64."""
    rng = np.random.default_rng(zlib.crc32(str(sizes).encode()))
    spec = rng_spec("mlp", rng, sizes=sizes)
    n = 5003
    pos = np.stack([rng.uniform(-0.2, 1.2, n) * (spec["max"] - spec["min"]) + spec["min"], rng.uniform(-2.5, 2.5, n)], 1)
    w = rng.uniform(0, 2, n)
    fr = biasspec.oracle_basis(spec, pos)
    from oracle import ves
    for prec, tol, dt in PREC:
        e = biasspec.make_ensemble(n, 2, 0, prec, potentials.DOUBLE_WELL_2D, None, spec)
        tpos = torch.tensor(pos, dtype=dt, device=DEV)
        e.set_state(tpos.T.contiguous(), None)
        p_dev = pos.astype(e.np_dtype).astype(np.float64)
        inside = biasspec.oracle_inside(spec, p_dev)
        for variant in (8, 6):
            e.set_tuning(variant)
            for domain in (False, True):
                e.set_moment_domain(domain)
                for weights in (None, w):
                    tws = None if weights is None else torch.tensor(weights, dtype=dt, device=DEV)
                    wm = (inside.astype(np.float64) if domain else np.ones(n)) * (1.0 if weights is None else weights)
                    sw_r, sf_r, _ = ves.moments(fr, wm)
                    sw, sf, _ = e.moments(row_weights=tws)
                    assert abs(float(sw.cpu()[0]) - sw_r) / sw_r < (1e-12 if prec == "fp64" and variant == 6 else 1e-6)
                    scale = np.maximum(np.abs(fr).T @ wm, 1.0)
                    err = np.max(np.abs(sf.cpu().numpy() - sf_r) / scale)
                    # FP32: a row within rounding of a switch point of the second layer can fall on the other side, and dV/dW2[j][i]
                    # jumps there by wout_j h_i -- for a parameter only a few rows feed, one such row is 2e-3 of the sum in the
                    # per-row FP32 kernel (measured, tools/diag/nn_moments_diag.py); the pieces come from FP64 breakpoints: 3e-5
                    bound = 1e-6 if prec == "fp64" else (tol if variant == 8 else 5e-3)
                    assert err < bound, (sizes, prec, variant, domain, err)
        e.close()


def test_moments_mlp_rejects_covariance():
    rng = np.random.default_rng(0)
    e = biasspec.make_ensemble(8, 2, 0, "fp32", potentials.DOUBLE_WELL_2D, None, rng_spec("mlp", rng, sizes=[4, 4]))
    with pytest.raises(NotImplementedError):
        e.moments(cov=True)
    e.set_bias_none()
    with pytest.raises(RuntimeError):
        e.moments()
    e.close()


def test_histogram_vs_numpy():
    rng = np.random.default_rng(2)
    n = 200003
    pos = np.stack([rng.normal(0, 1.2, n), rng.normal(0.3, 0.8, n)], 1)
    pos[:3] = [[2.25, 2.0], [-2.25, -2.0], [2.25, 0.0]]          # right edges land in the last bin
    w = rng.uniform(0, 1, n)
    for prec, tol, dt in PREC:
        e = biasspec.make_ensemble(n, 2, 0, prec, potentials.DOUBLE_WELL_2D, None, dict(kind="none"))
        p32 = pos.astype(e.np_dtype).astype(np.float64)      # bin what the device sees
        tpos = torch.tensor(pos, dtype=dt, device=DEV)
        e.set_state(tpos.T.contiguous(), None)
        ref2, _, _ = np.histogram2d(p32[:, 0], p32[:, 1], bins=[100, 100], range=[[-2.25, 2.25], [-2, 2]])
        c2 = e.histogram((-2.25, 2.25, -2, 2), (100, 100))
        assert np.array_equal(c2.cpu().numpy(), ref2)
        c2b = e.histogram((-2.25, 2.25, -2, 2), (100, 100), pos=pos.astype(e.np_dtype))
        assert np.array_equal(c2b, ref2)
        ref1, _ = np.histogram(p32[:, 1], bins=64, range=(-2, 2))
        c1 = e.histogram((-2, 2), 64, axis=1)
        assert np.array_equal(c1.cpu().numpy(), ref1)
        refw, _ = np.histogram(p32[:, 0], bins=200, range=(-2.25, 2.25), weights=w.astype(e.np_dtype).astype(np.float64))
        cw = e.histogram((-2.25, 2.25), 200, axis=0, row_weights=torch.tensor(w, dtype=dt, device=DEV))
        assert np.allclose(cw.cpu().numpy(), refw, rtol=1e-5)
        # accumulation into an existing array, and a bin count beyond the shared-memory path
        c1 = e.histogram((-2, 2), 64, axis=1, counts=c1)
        assert np.array_equal(c1.cpu().numpy(), 2 * ref1)
        big, _, _ = np.histogram2d(p32[:, 0], p32[:, 1], bins=[150, 150], range=[[-2.25, 2.25], [-2, 2]])
        assert np.array_equal(e.histogram((-2.25, 2.25, -2, 2), (150, 150)).cpu().numpy(), big)
        e.close()


def test_equilibrium_sampling_double_well_fes():
    """Unbiased ensemble in the single well x^2 + y^2 reproduces kT/2 per mode (size-independent property)."""
    e = biasspec.make_ensemble(200000, 2, 5, "fp32", potentials.SINGLE_WELL_2D, None, dict(kind="none"))
    k = langevin.constants()
    e.set_state(np.random.default_rng(0).standard_normal((2, 200000)).astype(np.float32) * np.sqrt(k["kT"] / 2), None)
    e.init_velocities()
    e.step(2000)
    x, v = e.get_state(host=True)
    assert abs(np.mean(x.astype(np.float64) ** 2) / (k["kT"] / 2) - 1) < 0.02
    assert abs(np.mean(v.astype(np.float64) ** 2) / k["kT"] - 1) < 0.02
    e.close()


@pytest.mark.parametrize("deg,sizes", [(9, (10007,)), (15, (31, 10007, 300000)), (19, (10007,)), (31, (77, 5003)), (63, (2003,))],
                         ids=["K100", "K256", "K400", "K1024", "K4096"])
def test_covariance_tensor_core_path(deg, sizes):
    """K > 64 covariance of an FP32 ensemble runs on tcgen05 (FP16 operand tiles: the 10-bit mantissa of TF32 for factors
    |f| <= 1; FP32 accumulate in TMEM), tiled into 256 x 256 output tiles (K = 400: padded tiles, K = 1024: 10
    upper-triangular tiles, K = 4096: all 136 tiles of the largest basis the ABI accepts): agrees with the FP64 oracle
    to the operand rounding (2^-11 per table entry and per product, averaging out over the rows) and with the CUDA-core
    kernel; symmetric exactly; row 0 reproduces the first moments (f_0 = 1).  Also the in-box domain (rows outside get
    weight 0) and row weights of either sign (sqrt(|w|) folded into the operands, negative rows in a second pass)."""
    from oracle import ves
    rng = np.random.default_rng(12)
    spec = rng_spec("leg2d", rng, deg_x=deg, deg_y=deg)
    for n in sizes:
        pos = np.stack([rng.uniform(-3, 3, n), rng.uniform(-2.5, 2.5, n)], 1)
        fr = biasspec.oracle_basis(spec, pos)
        sw_r, sf_r, sff_r = ves.moments(fr)
        e = biasspec.make_ensemble(n, 2, 0, "fp32", potentials.DOUBLE_WELL_2D, None, spec)
        e.set_state(torch.tensor(pos.T.copy(), dtype=torch.float32, device=DEV), None)
        sw, sf, sff = e.moments(cov=True)
        sff = sff.cpu().numpy()
        assert np.array_equal(sff, sff.T)
        scale = np.sqrt(np.outer(np.diag(sff_r), np.diag(sff_r)))          # |C_kl| <= sqrt(C_kk C_ll)
        assert np.max(np.abs(sff - sff_r) / scale) < 1e-3, n
        assert np.max(np.abs(sff[0] - sf_r) / np.maximum(np.abs(fr).sum(0), 1.0)) < 1e-3
        if deg <= 31:
            e.set_tuning(98)                                                # same call on the CUDA cores
            _, _, sff_cc = e.moments(cov=True)
            e.set_tuning(97)
            assert np.max(np.abs(sff_cc.cpu().numpy() - sff_r) / scale) < 1e-5
        if n >= 2000 and deg <= 31:
            # in-box domain: rows whose scaled coordinate clamps count with weight 0
            mm = spec["minmax"]
            inside = (pos[:, 0] >= mm[0]) & (pos[:, 0] <= mm[1]) & (pos[:, 1] >= mm[2]) & (pos[:, 1] <= mm[3])
            e.set_moment_domain(True)
            swi, _, sffi = e.moments(cov=True)
            e.set_moment_domain(False)
            _, _, ref_i = ves.moments(fr[inside])
            assert float(swi) == inside.sum()
            assert np.max(np.abs(sffi.cpu().numpy() - ref_i) / scale) < 1e-3
            # row weights: positive only (one pass), then of both signs (second pass for the negative rows)
            for w in (rng.uniform(0.0, 7.0, n), rng.standard_normal(n) * 3.0):
                _, _, ref_w = ves.moments(fr, w)
                wt = torch.tensor(w, dtype=torch.float32, device=DEV)
                rows = torch.tensor(pos, dtype=torch.float32, device=DEV)
                _, sfw, sffw = e.moments(pos=rows, row_weights=wt, cov=True)
                sffw = sffw.cpu().numpy()
                wscale = np.sqrt(np.outer(np.diag(ves.moments(fr, np.abs(w))[2]), np.diag(ves.moments(fr, np.abs(w))[2])))
                assert np.array_equal(sffw, sffw.T)
                assert np.max(np.abs(sffw - ref_w) / wscale) < 1e-3
        e.close()


def test_covariance_tensor_core_at_full_size():
    """10^7 rows (configs[2]'s ensemble on ONE GPU; 1.25 x 10^6 per GPU is the sharded case) at K = 256.  The tensor core
    truncates when it aligns the addends of its FP32 accumulation: measured (tools/diag/cov_accum_diag.py) the diagonal
    comes out low by 2^-27 per accumulated row, 4e-4 if a CTA kept its 67 000 rows in TMEM.  The kernel therefore drains
    the accumulators into its partial tile every 16384 rows (rounded FP32 adds) and the partials are summed in FP64:
    bias <= 1.1e-4 of sqrt(C_kk C_ll) whatever n is, off-diagonal entries 3e-6.  Checked against the CUDA-core kernel,
    which promotes to FP64 every 512 rows and is itself pinned to the FP64 oracle above.  Size-independent properties at
    full size: exact symmetry, C_00 = n, row 0 = first moments."""
    n = 10 ** 7
    rng = np.random.default_rng(3)
    spec = rng_spec("leg2d", rng, deg_x=15, deg_y=15)
    e = biasspec.make_ensemble(n, 2, 0, "fp32", potentials.DOUBLE_WELL_2D, None, spec)
    g = torch.Generator(device=DEV).manual_seed(5)
    pos = torch.empty((2, n), dtype=torch.float32, device=DEV)
    pos[0].uniform_(-2.5, 2.5, generator=g)
    pos[1].uniform_(-2.0, 2.0, generator=g)
    e.set_state(pos, None)
    sw, sf, sff = e.moments(cov=True)
    e.set_tuning(98)
    _, _, ref = e.moments(cov=True)
    e.set_tuning(97)
    sff, ref = sff.cpu().numpy(), ref.cpu().numpy()
    assert float(sw) == n and sff[0, 0] == n and np.array_equal(sff, sff.T)
    scale = np.sqrt(np.outer(np.diag(ref), np.diag(ref)))
    err = (sff - ref) / scale
    assert np.max(np.abs(err)) < 2e-4, np.max(np.abs(err))
    assert np.max(np.abs(err - np.diag(np.diag(err)))) < 2e-5          # entries of mixed-sign sums carry no truncation bias
    assert np.max(np.abs(sff[0] - sf.cpu().numpy())) < 2e-4 * n
    e.close()


@pytest.mark.gpu
def test_device_coefficients_on_concurrent_streams():
    """Handles whose coefficients were set from device memory (pyves_set_bias_legendre{1d,2d}_device) share ONE
    __constant__ weight array per device.  Independent handles with DIFFERENT coefficients, launched alternately on
    different streams so that their kernels overlap in time, must each see their own weights (the re-copy waits
    for the other stream's readers); a follower made with pyves_set_bias_like shares the leader's weights.  Every
    run equals, bit for bit, the same handle fed from the host."""
    import torch
    kT = 8.31446261815324e-3 * 300
    rng = np.random.default_rng(3)
    n, stride, rounds = 300_000, 150, 3

    def make(offset, count=n):
        e = _lib.Ensemble(count, dims=2, seed=5, precision="fp32")
        e.set_integrator(1.0, kT, 20.0, 0.01)
        e.set_potential(_lib.POT_DOUBLE_WELL_2D)
        e.set_walker_offset(offset)
        assert e.init_positions_mc((-2.5, 2.5, -2.0, 2.0), 1 / kT) == 0
        e.init_velocities()
        return e

    for kind in ("2d", "1d"):
        if kind == "2d":
            ws = [rng.normal(size=(16, 16)) * 0.5 / (1 + np.add.outer(np.arange(16), np.arange(16))) for _ in range(2)]
            set_host = lambda e, w: e.set_bias_legendre2d(15, 15, (-2.5, 2.5, -2.0, 2.0), w)
            set_dev = lambda e, w: e.set_bias_legendre2d_device(15, 15, (-2.5, 2.5, -2.0, 2.0), w)
        else:
            ws = [rng.normal(size=6) * 2.0 for _ in range(2)]
            set_host = lambda e, w: e.set_bias_legendre1d(0, 5, -2.5, 2.5, w)
            set_dev = lambda e, w: e.set_bias_legendre1d_device(0, 5, -2.5, 2.5, w)
        # reference: host-fed handles, one after the other
        want = []
        for k in range(2):
            e = make(k * n)
            set_host(e, ws[k])
            for _ in range(rounds):
                e.step(stride)
            want.append(e.get_state(host=True))
            e.close()
        # two independent device-fed handles on their own streams, launches interleaved
        hs = [make(k * n) for k in range(2)]
        wd = [torch.tensor(w, dtype=torch.float64, device="cuda").contiguous() for w in ws]
        streams = [torch.cuda.Stream() for _ in range(2)]
        torch.cuda.synchronize()
        for k in range(2):
            with torch.cuda.stream(streams[k]):
                set_dev(hs[k], wd[k])
        for _ in range(rounds):
            for k in range(2):
                with torch.cuda.stream(streams[k]):
                    hs[k].step(stride)
        torch.cuda.synchronize()
        for k in range(2):
            x, v = hs[k].get_state(host=True)
            assert np.array_equal(x, want[k][0]) and np.array_equal(v, want[k][1]), (kind, k)
        # a follower of handle 0 on a third stream, interleaved with handle 1: same walkers as handle 0 had
        f = make(0)
        s3 = torch.cuda.Stream()
        with torch.cuda.stream(s3):
            f.set_bias_like(hs[0])
        for _ in range(rounds):
            with torch.cuda.stream(streams[1]):
                hs[1].step(stride)
            with torch.cuda.stream(s3):
                f.step(stride)
        torch.cuda.synchronize()
        x, v = f.get_state(host=True)
        assert np.array_equal(x, want[0][0]) and np.array_equal(v, want[0][1]), kind
        # the same two handles driven from two host threads (ctypes releases the GIL during the calls)
        import threading
        gs = [make(k * n) for k in range(2)]
        errs = []

        def drive(k):
            try:
                torch.cuda.set_device(0)
                with torch.cuda.stream(streams[k]):
                    set_dev(gs[k], wd[k])
                    for _ in range(rounds):
                        gs[k].step(stride)
            except Exception as exc:       # noqa: BLE001
                errs.append(exc)

        ths = [threading.Thread(target=drive, args=(k,)) for k in range(2)]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        torch.cuda.synchronize()
        assert not errs, errs
        for k in range(2):
            x, v = gs[k].get_state(host=True)
            assert np.array_equal(x, want[k][0]) and np.array_equal(v, want[k][1]), (kind, "threads", k)
        for e in hs + gs + [f]:
            e.close()
    # mismatched handles are refused
    a = _lib.Ensemble(8, dims=2, seed=1, precision="fp32")
    b = _lib.Ensemble(8, dims=2, seed=1, precision="fp64")
    a.set_bias_legendre1d(0, 5, -2.5, 2.5, np.zeros(6))
    with pytest.raises(Exception):
        b.set_bias_like(a)           # precision differs
    c = _lib.Ensemble(8, dims=2, seed=1, precision="fp32")
    with pytest.raises(Exception):
        c.set_bias_like(a)           # the leader's coefficients are host-fed
    for e in (a, b, c):
        e.close()


@pytest.mark.gpu
def test_philox_sequence_across_call_index_word():
    """The light kernels advance Philox incrementally within a launch (PhiloxSeq), valid while the upper word of
    the call index (step / 2) is fixed; the library cuts launches at multiples of 2^33 steps.  One launch across
    that boundary equals single-step launches (each of which starts from its own counter), bit for bit, and the
    oracle's noise for those steps."""
    kT = 8.31446261815324e-3 * 300
    n, before, after = 1000, 7, 8
    start = (1 << 33) - before
    w = np.linspace(-1.0, 1.0, 6)
    runs = []
    for mode in ("one_launch", "single_steps"):
        e = _lib.Ensemble(n, dims=2, seed=99, precision="fp32")
        e.set_integrator(1.0, kT, 20.0, 0.01)
        e.set_potential(_lib.POT_DOUBLE_WELL_2D)
        e.set_bias_legendre1d(0, 5, -2.5, 2.5, w)
        assert e.init_positions_mc((-2.5, 2.5, -2.0, 2.0), 1 / kT) == 0
        e.init_velocities()
        e.step_counter = start
        if mode == "one_launch":
            e.step(before + after)
        else:
            for _ in range(before + after):
                e.step(1)
        assert e.step_counter == start + before + after
        runs.append(e.get_state(host=True))
        e.close()
    assert np.array_equal(runs[0][0], runs[1][0]) and np.array_equal(runs[0][1], runs[1][1])
    # and against the oracle's Philox stream for the same steps (FP64 handle; the call index crosses 2^32)
    rng = np.random.default_rng(2)
    spec = dict(kind="leg1d", axis='x', degree=5, min=-2.5, max=2.5, weights=w)
    k = langevin.constants()
    m = 64
    x0 = np.stack([rng.uniform(-1.6, 1.5, m), rng.uniform(-0.3, 0.3, m)], 1)
    v0 = rng.standard_normal((m, 2)) * k["sigma_v"]
    xr, vr = langevin.run_philox(x0, v0, biasspec.total_force_fn(potentials.DOUBLE_WELL_2D, None, spec), k, 99, 3, start, before + after)
    e = biasspec.make_ensemble(m, 2, 99, "fp64", potentials.DOUBLE_WELL_2D, None, spec)
    e.set_walker_offset(3)
    e.set_state(torch.tensor(x0.T.copy(), dtype=torch.float64, device=DEV), torch.tensor(v0.T.copy(), dtype=torch.float64, device=DEV))
    e.step_counter = start
    e.step(before + after)
    xg, vg = e.get_state(host=True)
    e.close()
    assert rel_err(xg.T, xr) < 1e-9 and rel_err(vg.T, vr) < 1e-9
