"""Oracle (test infrastructure): Philox4x32-10 and the uniform->normal mapping.

Algorithm: Salmon, Moraes, Dror, Shaw, "Parallel random numbers: as easy as 1, 2, 3",
SC'11 (Random123).  Multipliers 0xD2511F53 / 0xCD9E8D57, Weyl key bumps
0x9E3779B9 / 0xBB67AE85, 10 rounds.  Known-answer vectors: SURVEY Appendix A.2.

The reference draws its noise inside OpenMM (``ves/langevin_dynamics.py:49-53,188``), whose
generator is not reproducible outside OpenMM; the stream below is the new framework's own
specification (DESIGN.md "Noise"), shared by the CUDA kernels and this oracle.
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK32 = np.uint64(0xFFFFFFFF)

# stream ids (top byte of counter word 3)
STREAM_DYN2 = 0   # dynamics, 2 simulated dimensions: one call feeds two steps
STREAM_DYN3 = 1   # dynamics, 3 simulated dimensions: one call per step
STREAM_VEL = 2    # Maxwell-Boltzmann velocities
STREAM_POS = 3    # Monte-Carlo initial placement


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10.  Counter words are uint32 arrays (broadcastable); key words ints."""
    c0 = np.asarray(c0, dtype=np.uint64) & MASK32
    c1 = np.asarray(c1, dtype=np.uint64) & MASK32
    c2 = np.asarray(c2, dtype=np.uint64) & MASK32
    c3 = np.asarray(c3, dtype=np.uint64) & MASK32
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK32
        n0 = hi1 ^ c1 ^ np.uint64(k0)
        n2 = hi0 ^ c3 ^ np.uint64(k1)
        c0, c1, c2, c3 = n0, lo1, n2, lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return (c0.astype(np.uint32), c1.astype(np.uint32), c2.astype(np.uint32), c3.astype(np.uint32))


def seed_key(seed):
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return seed & 0xFFFFFFFF, seed >> 32


def counter(gid, call, stream):
    """Counter layout: (gid_lo, gid_hi, call_lo, (call_hi & 0xFFFFFF) | stream << 24)."""
    gid = np.asarray(gid, dtype=np.uint64)
    call = np.uint64(call)
    c0 = gid & MASK32
    c1 = gid >> np.uint64(32)
    c2 = call & MASK32
    c3 = ((call >> np.uint64(32)) & np.uint64(0xFFFFFF)) | (np.uint64(stream) << np.uint64(24))
    return c0, c1, c2, c3


def uniform(r):
    """(r + 0.5) * 2^-32 in (0, 1)."""
    return (np.asarray(r, dtype=np.float64) + 0.5) * 2.0 ** -32


def box_muller(r0, r1):
    """Two uint32 -> two N(0,1): rad = sqrt(-2 ln u1), theta = pi (2 u2 - 1); (rad cos, rad sin)."""
    u1 = uniform(r0)
    u2 = uniform(r1)
    rad = np.sqrt(-2.0 * np.log(u1))
    theta = np.pi * (2.0 * u2 - 1.0)
    return rad * np.cos(theta), rad * np.sin(theta)


def dynamics_noise(seed, gids, step, dims=2):
    """Noise of global step ``step`` for walkers ``gids`` -> array [len(gids), dims]."""
    k0, k1 = seed_key(seed)
    gids = np.asarray(gids, dtype=np.uint64)
    if dims == 2:
        o = philox4x32_10(*counter(gids, step >> 1, STREAM_DYN2), k0, k1)
        if step & 1:
            zx, zy = box_muller(o[2], o[3])
        else:
            zx, zy = box_muller(o[0], o[1])
        return np.stack([zx, zy], axis=1)
    if dims == 3:
        o = philox4x32_10(*counter(gids, step, STREAM_DYN3), k0, k1)
        zx, zy = box_muller(o[0], o[1])
        zz, _ = box_muller(o[2], o[3])
        return np.stack([zx, zy, zz], axis=1)
    raise ValueError("dims must be 2 or 3")


def velocity_noise(seed, gids):
    """Three N(0,1) per walker for the Maxwell-Boltzmann draw -> [len(gids), 3]."""
    k0, k1 = seed_key(seed)
    o = philox4x32_10(*counter(np.asarray(gids, dtype=np.uint64), 0, STREAM_VEL), k0, k1)
    zx, zy = box_muller(o[0], o[1])
    zz, _ = box_muller(o[2], o[3])
    return np.stack([zx, zy, zz], axis=1)


def placement_uniforms(seed, gids, trial):
    """Two U(0,1) per walker for Monte-Carlo placement trial ``trial`` -> (ux, uy)."""
    k0, k1 = seed_key(seed)
    o = philox4x32_10(*counter(np.asarray(gids, dtype=np.uint64), trial, STREAM_POS), k0, k1)
    return uniform(o[0]), uniform(o[1])
