"""Oracle (test infrastructure): coefficient-update rules.

1. Reference-literal loss of ``VESBias_SingleParticle_1D.update`` (``ves/bias.py:206-253``):
   ``L = (1/beta) log sum_{t in window} exp(beta V(x_t)) + sum_i p_i V(x_i^target)``
   with the window ``max(T - update_steps, 0) .. T`` (``:219``) and the target nodes fed
   through the module as raw x coordinates (``:229-233``).  PINNED by golden vectors.
2. Textbook VES (Valsson & Parrinello, PRL 113, 090601 (2014)): gradient
   ``-<f>_V + <f>_p``, Hessian ``beta Cov_V(f)``, Bach-Moulines averaged SGD; well-tempered
   target (Valsson & Parrinello, JCTC 11, 1996 (2015)).  PARITY UNPINNED -- the reference
   only names these options (``README.md:10-31``, ``ves/bias.py:65``).
"""
import numpy as np


# ----------------------------------------------------------------------------- reference-literal
def literal_loss_and_grad(V_traj, dV_traj, V_tgt, dV_tgt, p_tgt, beta):
    """V_* are energies [T] / [M]; dV_* their parameter gradients [T, P] / [M, P].

    Returns (loss, dloss/dtheta[P])."""
    V_traj = np.asarray(V_traj, dtype=np.float64)
    m = (beta * V_traj).max()
    e = np.exp(beta * V_traj - m)
    lse = m + np.log(e.sum())
    loss = lse / beta + float(np.dot(p_tgt, V_tgt))
    sm = e / e.sum()
    grad = sm @ np.asarray(dV_traj) + np.asarray(p_tgt) @ np.asarray(dV_tgt)
    return loss, grad


def window(T, update_steps):
    return max(T - update_steps, 0), T


def sgd_step(theta, grad, lr, momentum=0.0, state=None):
    """torch.optim.SGD (no dampening / nesterov / weight decay)."""
    if momentum:
        buf = grad.copy() if state is None else momentum * state + grad
        return theta - lr * buf, buf
    return theta - lr * grad, None


def adam_step(theta, grad, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, state=None):
    """torch.optim.Adam, one step; state = (step, m, v)."""
    t, m, v = (0, np.zeros_like(theta), np.zeros_like(theta)) if state is None else state
    t += 1
    m = betas[0] * m + (1 - betas[0]) * grad
    v = betas[1] * v + (1 - betas[1]) * grad * grad
    mhat = m / (1 - betas[0] ** t)
    denom = np.sqrt(v) / np.sqrt(1 - betas[1] ** t) + eps
    return theta - lr * mhat / denom, (t, m, v)


def rmsprop_step(theta, grad, lr=1e-2, alpha=0.99, eps=1e-8, state=None):
    """torch.optim.RMSprop (no momentum, not centered); state = square average."""
    sq = np.zeros_like(theta) if state is None else state
    sq = alpha * sq + (1 - alpha) * grad * grad
    return theta - lr * grad / (np.sqrt(sq) + eps), sq


# ----------------------------------------------------------------------------- textbook VES
def moments(f, w=None):
    """sum_w, sum_w f [K], sum_w f f^T [K, K] from basis values f [N, K]."""
    f = np.asarray(f, dtype=np.float64)
    w = np.ones(f.shape[0]) if w is None else np.asarray(w, dtype=np.float64)
    return w.sum(), w @ f, (f * w[:, None]).T @ f


def gradient_hessian(sum_w, sum_f, sum_ff, f_target, beta):
    mean = sum_f / sum_w
    grad = -mean + f_target
    hess = None
    if sum_ff is not None:
        hess = beta * (sum_ff / sum_w - np.outer(mean, mean))
    return grad, hess


class AveragedSGD:
    """alpha^{n+1} = alpha^n - mu [grad(abar^n) + H(abar^n)(alpha^n - abar^n)];
    abar^n = mean of alpha^0..alpha^n (``decay=None``) or the exponentially weighted mean
    abar <- decay abar + (1 - decay) alpha; the applied bias uses abar."""

    def __init__(self, alpha0, mu, decay=None):
        self.alpha = np.array(alpha0, dtype=np.float64)
        self.abar = self.alpha.copy()
        self.n = 0
        self.mu = mu
        self.decay = decay

    def step(self, grad, hess=None):
        d = grad.copy()
        if hess is not None:
            d = d + hess @ (self.alpha - self.abar)
        self.alpha = self.alpha - self.mu * d
        self.n += 1
        if self.decay is None:
            self.abar = self.abar + (self.alpha - self.abar) / (self.n + 1)
        else:
            self.abar = self.decay * self.abar + (1 - self.decay) * self.alpha
        return self.abar


def target_average(f_grid, p_grid):
    """<f_k>_p by quadrature with normalised node weights: f_grid [M, K], p_grid [M]."""
    p = np.asarray(p_grid, dtype=np.float64)
    return (p / p.sum()) @ np.asarray(f_grid, dtype=np.float64)


def well_tempered_target(V_grid, p_grid, beta, bias_factor):
    """One iteration p' ~ exp(-(beta/gamma) F), F = -V - (1/beta) ln p (normalised to sum 1)."""
    p = np.asarray(p_grid, dtype=np.float64)
    F = -np.asarray(V_grid, dtype=np.float64) - np.log(np.maximum(p, 1e-300)) / beta
    F = F - F.min()
    q = np.exp(-(beta / bias_factor) * F)
    return q / q.sum()


def fes_from_bias(V_grid, p_grid, beta):
    """F(s) = -V(s) - (1/beta) ln p(s), shifted to min 0 (in kJ/mol)."""
    F = -np.asarray(V_grid, dtype=np.float64) - np.log(np.maximum(p_grid, 1e-300)) / beta
    return F - F.min()
