"""Oracle (test infrastructure): the NNBasis1D multilayer perceptron.

Layer structure follows ``ves/basis.py:184-205``: ``Linear(1, h0)`` with NO activation, then
``Linear(h_{l-1}, h_l)`` + act for l >= 1, then ``Linear(h_last, 1)``.  Input scaling
``s = (q - min) / (max - min)`` follows ``:227`` (no clamp).  The reference gets the force
from autograd through TorchForce (``ves/bias.py:58``); here the derivative is propagated in
forward mode (the input is a scalar).  ReLU'(0) = 0 as in torch.

Parameters are one flat FP64 vector in torch ``parameters()`` order:
W0[h0,1], b0[h0], W1[h1,h0], b1[h1], ..., Wout[1,h_last], bout[1].
"""
import numpy as np

ACT_RELU = 0
ACT_TANH = 1


def n_params(sizes):
    dims = [1] + list(sizes) + [1]
    return sum(dims[i] * dims[i + 1] + dims[i + 1] for i in range(len(dims) - 1))


def unpack(theta, sizes):
    dims = [1] + list(sizes) + [1]
    theta = np.asarray(theta, dtype=np.float64)
    out = []
    o = 0
    for i in range(len(dims) - 1):
        W = theta[o:o + dims[i] * dims[i + 1]].reshape(dims[i + 1], dims[i])
        o += W.size
        b = theta[o:o + dims[i + 1]]
        o += b.size
        out.append((W, b))
    assert o == theta.size
    return out


def _act(z, act):
    if act == ACT_RELU:
        return np.maximum(z, 0.0), (z > 0).astype(np.float64)
    if act == ACT_TANH:
        t = np.tanh(z)
        return t, 1.0 - t * t
    raise ValueError("unknown activation")


def forward(s, theta, sizes, act=ACT_RELU):
    """V(s) and dV/ds for scalar inputs s [N]."""
    layers = unpack(theta, sizes)
    s = np.asarray(s, dtype=np.float64)
    h = s[:, None] * layers[0][0][:, 0][None, :] + layers[0][1][None, :]   # no activation
    dh = np.broadcast_to(layers[0][0][:, 0][None, :], h.shape).copy()
    for (W, b) in layers[1:-1]:
        z = h @ W.T + b
        dz = dh @ W.T
        h, g = _act(z, act)
        dh = dz * g
    W, b = layers[-1]
    return (h @ W.T + b)[:, 0], (dh @ W.T)[:, 0]


def nnbasis1d(pos, lo, hi, axis, theta, sizes, act=ACT_RELU):
    """Energy [N] and force [N, D] of NNBasis1D(min=lo, max=hi, axis, hidden_layer_sizes=sizes)."""
    pos = np.asarray(pos, dtype=np.float64)
    a = {'x': 0, 'y': 1}[axis]
    s = (pos[:, a] - lo) / (hi - lo)
    V, dV = forward(s, theta, sizes, act)
    F = np.zeros_like(pos)
    F[:, a] = -dV / (hi - lo)
    return V, F


def param_grad(s, theta, sizes, act=ACT_RELU, row_weights=None):
    """sum_rows g_row dV(s_row)/dtheta (flat, same packing as theta); g defaults to ones."""
    layers = unpack(theta, sizes)
    s = np.asarray(s, dtype=np.float64)
    g_out = np.ones_like(s) if row_weights is None else np.asarray(row_weights, dtype=np.float64)
    hs = [s[:, None]]
    h = hs[0] @ layers[0][0].T + layers[0][1]
    hs.append(h)
    gates = []
    for (W, b) in layers[1:-1]:
        z = h @ W.T + b
        h, g = _act(z, act)
        gates.append(g)
        hs.append(h)
    grads = [None] * len(layers)
    W, b = layers[-1]
    delta = g_out[:, None] * np.ones((s.shape[0], 1))
    grads[-1] = (delta.T @ hs[-1], delta.sum(0))
    delta = delta @ W
    for li in range(len(layers) - 2, 0, -1):
        delta = delta * gates[li - 1]
        grads[li] = (delta.T @ hs[li], delta.sum(0))
        delta = delta @ layers[li][0]
    grads[0] = (delta.T @ hs[0], delta.sum(0))
    return np.concatenate([np.concatenate([gw.ravel(), gb.ravel()]) for gw, gb in grads])
