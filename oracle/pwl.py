"""Oracle (test infrastructure): a ReLU NNBasis1D as the piecewise-linear function of its one input that it is.

Not an algorithm of the reference (``ves/basis.py:184-232`` evaluates the network through torch for every particle); it restates in numpy
the construction the CUDA side uses (``mlp_pwl_build_kernel`` / ``BiasPwl1D``, ``pwl_hist_kernel``) so that its two claims can be checked on
the CPU against ``oracle/mlp.py``:

* dV/ds is piecewise CONSTANT: for any depth, the pre-activations of a layer are linear in s on an interval where the activation pattern of
  all earlier layers is fixed, so every unit adds at most one switch point per such interval;
* on a piece every parameter gradient dV/dtheta is AFFINE in s, so a weighted sum of gradients over rows equals the sum over the pieces
  of n_p dV/dtheta(mean_p).

``compile_pwl`` works for any number of hidden layers (the CUDA kernel handles two and three) by refining the intervals layer by layer.
"""
import numpy as np

from . import mlp


def compile_pwl(theta, sizes):
    """Sorted breakpoints brk [P] and slopes [P + 1] of dV/ds: slope[k] holds for brk[k-1] < s < brk[k]."""
    layers = mlp.unpack(theta, sizes)
    # affine maps of the current layer's OUTPUT on every interval: h = A s + B, intervals given by the sorted cuts
    cuts = np.zeros(0)
    A = [layers[0][0][:, 0].copy()]               # first Linear has no activation (ves/basis.py:197-202)
    B = [layers[0][1].copy()]
    for (W, b) in layers[1:-1]:
        new_cuts, newA, newB = [], [], []
        edges = np.concatenate([[-np.inf], cuts, [np.inf]])
        for r in range(len(edges) - 1):
            lo, hi = edges[r], edges[r + 1]
            za, zb = W @ A[r], W @ B[r] + b       # pre-activations z = za s + zb on this interval
            with np.errstate(divide="ignore", invalid="ignore"):
                kap = np.where(za != 0.0, -zb / za, np.nan)
            inside = np.sort(kap[(kap > lo) & (kap < hi)])
            sub = np.concatenate([[lo], inside, [hi]])
            for q in range(len(sub) - 1):
                a, c = sub[q], sub[q + 1]
                mid = 0.5 * (a + c) if np.isfinite(a) and np.isfinite(c) else (c - 1.0 if np.isfinite(c) else (a + 1.0 if np.isfinite(a) else 0.0))
                m = (za * mid + zb > 0).astype(np.float64)      # active set evaluated inside the piece, never inferred
                newA.append(za * m)
                newB.append(zb * m)
                if q + 1 < len(sub) - 1 or r + 1 < len(edges) - 1:
                    new_cuts.append(c)
        cuts, A, B = np.asarray(new_cuts, dtype=np.float64), newA, newB
    Wout = layers[-1][0][0]
    slope = np.array([Wout @ a for a in A])
    return cuts, slope


def slope_at(s, brk, slope):
    return slope[np.searchsorted(brk, np.asarray(s, dtype=np.float64), side="left")]


def param_grad_through_pieces(s, theta, sizes, row_weights=None):
    """sum_r w_r dV(s_r)/dtheta computed from ONE pseudo row per piece: (mean of s over the rows of the piece, sum of their weights)."""
    s = np.asarray(s, dtype=np.float64)
    w = np.ones_like(s) if row_weights is None else np.asarray(row_weights, dtype=np.float64)
    brk, _ = compile_pwl(theta, sizes)
    piece = np.searchsorted(brk, s, side="left")
    n_p = np.bincount(piece, weights=w, minlength=len(brk) + 1)
    s_p = np.bincount(piece, weights=w * s, minlength=len(brk) + 1)
    keep = n_p != 0
    return mlp.param_grad(s_p[keep] / n_p[keep], theta, sizes, mlp.ACT_RELU, n_p[keep]), int(keep.sum())
