"""Oracle (test infrastructure): the two rewrites behind the tensor-core Legendre step kernel (``csrc/langevin_tc.cu``), in numpy.

Neither is an algorithm of the reference (``ves/basis.py:86-156`` evaluates the recursion and lets autograd differentiate it); they are
restated here so that the CPU suite can check them against ``oracle/legendre.py``:

* folded derivatives: with ``P'_n = sum_{k = n-1, n-3, ...} (2k+1) P_k`` both force components of the tensor-product basis are
  contractions with Legendre VALUES only,
      dV/dsx = sum_k P_k(sx) A'_k,  A'_k = sum_j WA[k][j] P_j(sy),  WA[k][j] = (2k+1) sum_{i>k, i-k odd} w_ij
      dV/dsy = sum_i P_i(sx) B_i,   B_i  = sum_k WB[i][k] P_k(sy),  WB[i][k] = (2k+1) sum_{j>k, j-k odd} w_ij;
* scaled monic recursion: ``q_j = P_j 2^e_j / k_j`` (k_j the leading coefficient of P_j, e_j = round(log2 k_j)) obeys
  ``q_{j+1} = 2^d_j s q_j - cc_j q_{j-1}`` with d_j in {0, 1}, and ``P_j = kappa_j q_j`` with kappa_j in [0.71, 1.41].
"""
import numpy as np


def folded_matrices(w):
    """WA [Kx, Ky], WB [Kx, Ky] of the coefficient matrix w [Kx, Ky]."""
    w = np.asarray(w, dtype=np.float64)
    kx, ky = w.shape
    WA = np.zeros_like(w)
    WB = np.zeros_like(w)
    for k in range(kx):
        WA[k] = (2 * k + 1) * w[k + 1::2].sum(axis=0)
    for k in range(ky):
        WB[:, k] = (2 * k + 1) * w[:, k + 1::2].sum(axis=1)
    return WA, WB


def monic_table(n):
    """cc [n-1], d [n-1], kappa [n] of the scaled monic recursion (the constexpr table make_monic_tab of the kernel)."""
    cc, d, kappa = np.zeros(n - 1), np.zeros(n - 1, dtype=int), np.ones(n)
    kap, e_prev, e_cur = 1.0, 0, 0
    for j in range(n - 1):
        kn = kap * (2 * j + 1) / (j + 1)
        dn = 0
        if kn >= np.sqrt(2.0):
            kn *= 0.5
            dn = 1
        e_next = e_cur + dn
        c = 0.0 if j == 0 else j * j / (4.0 * j * j - 1.0)
        cc[j], d[j], kappa[j + 1] = c * 2.0 ** (e_next - e_prev), dn, kn
        kap, e_prev, e_cur = kn, e_cur, e_next
    return cc, d, kappa


def monic_values(s, n):
    """q_j(s), j < n -> [n, len(s)]."""
    s = np.asarray(s, dtype=np.float64)
    cc, d, _ = monic_table(n)
    q = np.zeros((n,) + s.shape)
    q[0] = 1.0
    if n > 1:
        q[1] = s
    for j in range(1, n - 1):
        q[j + 1] = (2.0 * s if d[j] else s) * q[j] - cc[j] * q[j - 1]
    return q


def gradient_folded(sx, sy, w):
    """(dV/dsx, dV/dsy) from Legendre values only: the scaled monic values against kappa-scaled WA over WB, as the kernel contracts them."""
    w = np.asarray(w, dtype=np.float64)
    kx, ky = w.shape
    WA, WB = folded_matrices(w)
    _, _, kap_x = monic_table(max(kx, 2))
    _, _, kap_y = monic_table(max(ky, 2))
    qx, qy = monic_values(sx, max(kx, 2))[:kx], monic_values(sy, max(ky, 2))[:ky]
    scale = np.outer(kap_x[:kx], kap_y[:ky])
    A = (WA * scale) @ qy              # [Kx, N]
    B = (WB * scale) @ qy
    return np.sum(qx * A, axis=0), np.sum(qx * B, axis=0)
