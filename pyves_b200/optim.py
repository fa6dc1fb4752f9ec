"""Coefficient optimisers of the ensemble VES update (host side, FP64, K-sized vectors).

Textbook VES (Valsson & Parrinello, PRL 113, 090601 (2014)): with V(s; alpha) = sum_k alpha_k f_k(s),
    dOmega/dalpha_k = -<f_k>_V + <f_k>_p          H_kl = beta Cov_V(f_k, f_l)
and the averaged stochastic gradient descent of Bach & Moulines used by VES,
    alpha^{n+1} = alpha^n - mu [ grad(abar^n) + H(abar^n) (alpha^n - abar^n) ],   abar^n = mean(alpha^0..alpha^n),
with the applied bias built from abar.  The reference only names these options
(``README.md:22-31``; ``ves/bias.py:65`` "TODO: Implement analytical VES"), so parity for this module
is against ``oracle/ves.py`` (literature restatement), not against reference code.
"""
import numpy as np


class EmptyEnsembleAverage(ValueError):
    """The ensemble average has no sample (sum of weights <= 0)."""


def gradient_hessian(sum_w, sum_wf, sum_wff, f_target, beta):
    """(gradient [K], Hessian [K, K] or None) from ensemble moments and target averages.  ``sum_w <= 0`` (no
    sample in the averaging domain, e.g. every walker outside the basis box) raises ``EmptyEnsembleAverage``."""
    if not float(sum_w) > 0.0:
        raise EmptyEnsembleAverage("sum of the row weights is %r: no walker inside the averaging domain" % float(sum_w))
    mean = np.asarray(sum_wf, dtype=np.float64) / float(sum_w)
    grad = np.asarray(f_target, dtype=np.float64) - mean
    hess = None
    if sum_wff is not None:
        hess = beta * (np.asarray(sum_wff, dtype=np.float64) / float(sum_w) - np.outer(mean, mean))
    return grad, hess


class AveragedSGD:
    """Bach-Moulines averaged SGD.

    ``second_order`` adds the Hessian correction ``H (alpha - abar)``, i.e. the gradient is Taylor
    expanded from the averaged coefficients (where the ensemble is sampled) to the instantaneous
    ones; without it the iterates overshoot by the lag of the average and the scheme oscillates.
    ``averaging='running'`` is the plain mean of all iterates (the textbook choice, converging like
    1/n); ``averaging='ewma'`` is the exponentially weighted mean with factor ``decay`` (the
    "weighted average" option named in the reference's README, ``README.md:16-20``), which forgets
    the transient and is what the ensemble runs use.
    """

    def __init__(self, alpha0, mu=1.0, second_order=False, averaging='running', decay=0.9, scale=None):
        if averaging not in ('running', 'ewma'):
            raise ValueError("averaging must be 'running' or 'ewma'")
        self.alpha = np.array(alpha0, dtype=np.float64).ravel()
        self.abar = self.alpha.copy()
        self.mu = float(mu)
        self.second_order = bool(second_order)
        self.averaging = averaging
        self.decay = float(decay)
        # optional diagonal preconditioner: per-coefficient step sizes mu * scale_k.  For Legendre bases
        # scale_k = (2i+1)(2j+1) = 1 / Var_uniform(f_k) makes every mode relax at the same rate once the
        # biased ensemble is close to the (flat) target.
        self.scale = None if scale is None else np.array(scale, dtype=np.float64).ravel()
        self.n = 0

    def step(self, grad, hess=None):
        d = np.array(grad, dtype=np.float64)
        if self.second_order and hess is not None:
            d = d + hess @ (self.alpha - self.abar)
        if self.scale is not None:
            d = d * self.scale
        self.alpha = self.alpha - self.mu * d
        self.n += 1
        if self.averaging == 'ewma':
            self.abar = self.decay * self.abar + (1.0 - self.decay) * self.alpha
        else:
            self.abar = self.abar + (self.alpha - self.abar) / (self.n + 1)
        return self.abar

    def state_dict(self):
        return dict(alpha=self.alpha.copy(), abar=self.abar.copy(), mu=self.mu, second_order=self.second_order,
                    averaging=self.averaging, decay=self.decay, n=self.n)

    def load_state_dict(self, d):
        if len(d["alpha"]) != len(self.alpha):
            raise ValueError("checkpointed coefficients do not match the basis size")
        self.alpha = np.array(d["alpha"], dtype=np.float64)
        self.abar = np.array(d["abar"], dtype=np.float64)
        self.mu, self.second_order, self.n = float(d["mu"]), bool(d["second_order"]), int(d["n"])
        self.averaging, self.decay = d.get("averaging", "running"), float(d.get("decay", 0.9))


class RunningAverage:
    """Running mean of gradients over updates (the reference's README names running / weighted /
    moving-window averaging, ``README.md:16-20``); ``decay`` in (0, 1) gives the exponentially
    weighted variant, ``decay=None`` the plain running mean."""

    def __init__(self, decay=None):
        self.decay = decay
        self.value = None
        self.n = 0

    def update(self, g):
        g = np.asarray(g, dtype=np.float64)
        self.n += 1
        if self.value is None:
            self.value = g.copy()
        elif self.decay is None:
            self.value += (g - self.value) / self.n
        else:
            self.value = self.decay * self.value + (1 - self.decay) * g
        return self.value
