"""Oracle (test infrastructure): ctypes front end of the plain-C restatement ``oracle/c``.

``oracle/c/pyves_oracle.c`` is the CPU twin of the step kernel compiled with gcc + OpenMP; bench.py
uses it as the ``cpu_baseline`` ("port") and as the ``--impl reference`` arm, tests use it as a
faster checker.  It is validated against the numpy oracle in ``tests/test_oracle_c.py``.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "libpyves_oracle.so")

POT = {"double_well_2d": 3, "slip_bond_2d": 4, "szabo_berezhkovskii": 6}
BIAS = {"none": 0, "leg1d": 1, "leg2d": 2, "mlp": 3}


class _Cfg(ctypes.Structure):
    _fields_ = [("pot_kind", ctypes.c_int), ("pot", ctypes.c_double * 11), ("bias_kind", ctypes.c_int),
                ("axis", ctypes.c_int), ("deg_x", ctypes.c_int), ("deg_y", ctypes.c_int),
                ("lo_x", ctypes.c_double), ("hi_x", ctypes.c_double), ("lo_y", ctypes.c_double), ("hi_y", ctypes.c_double),
                ("w", ctypes.POINTER(ctypes.c_double)), ("n_hidden", ctypes.c_int), ("act", ctypes.c_int),
                ("sizes", ctypes.POINTER(ctypes.c_int)), ("theta", ctypes.POINTER(ctypes.c_double)),
                ("a", ctypes.c_double), ("b_over_m", ctypes.c_double), ("c_over_sqrtm", ctypes.c_double), ("dt", ctypes.c_double),
                ("seed", ctypes.c_uint64), ("gid0", ctypes.c_uint64)]


def build(force=False):
    src = os.path.join(_HERE, "c", "pyves_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-B", "-C", os.path.join(_HERE, "c")], check=True, capture_output=True)
    return LIB_PATH


_lib = None


def load():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.oracle_langevin2d.argtypes = [ctypes.POINTER(_Cfg), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                           ctypes.c_int64, ctypes.c_int64]
        _lib.oracle_moments.argtypes = [ctypes.POINTER(_Cfg), ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]
        _lib.oracle_moments_par.argtypes = [ctypes.POINTER(_Cfg), ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    return _lib


class Model:
    """Potential + bias + integrator constants for the C port (keeps the numpy buffers alive)."""

    def __init__(self, pot_kind, pot_params, spec, consts, seed=0, gid0=0):
        c = _Cfg()
        c.pot_kind = pot_kind
        for i, v in enumerate(pot_params or ()):
            c.pot[i] = v
        kind = spec["kind"]
        c.bias_kind = BIAS[kind]
        self._keep = []
        if kind == "leg1d":
            c.axis = {'x': 0, 'y': 1}[spec["axis"]]
            c.deg_x = spec["degree"]
            c.lo_x, c.hi_x = spec["min"], spec["max"]
            w = np.ascontiguousarray(spec["weights"], dtype=np.float64)
            self._keep.append(w)
            c.w = w.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
        elif kind == "leg2d":
            c.deg_x, c.deg_y = spec["deg_x"], spec["deg_y"]
            c.lo_x, c.hi_x, c.lo_y, c.hi_y = spec["minmax"]
            w = np.ascontiguousarray(spec["weights"], dtype=np.float64).ravel()
            self._keep.append(w)
            c.w = w.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
        elif kind == "mlp":
            c.axis = {'x': 0, 'y': 1}[spec["axis"]]
            c.lo_x, c.hi_x = spec["min"], spec["max"]
            s = np.ascontiguousarray(spec["sizes"], dtype=np.int32)
            th = np.ascontiguousarray(spec["theta"], dtype=np.float64)
            self._keep += [s, th]
            c.n_hidden, c.act = len(s), spec.get("act", 0)
            c.sizes = s.ctypes.data_as(ctypes.POINTER(ctypes.c_int))
            c.theta = th.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
        c.a, c.b_over_m, c.c_over_sqrtm, c.dt = consts["a"], consts["b_over_m"], consts["c_over_sqrtm"], consts["dt"]
        c.seed, c.gid0 = seed, gid0
        self.cfg = c

    def run(self, pos, vel, step0, nsteps):
        """pos, vel: float64 C-contiguous [2, n], updated in place."""
        assert pos.dtype == np.float64 and pos.flags.c_contiguous and vel.dtype == np.float64 and vel.flags.c_contiguous
        load().oracle_langevin2d(ctypes.byref(self.cfg), pos.ctypes.data, vel.ctypes.data, pos.shape[1], step0, nsteps)

    def moments(self, pos):
        K = (self.cfg.deg_x + 1) * ((self.cfg.deg_y + 1) if self.cfg.bias_kind == 2 else 1)
        out = np.zeros(K)
        load().oracle_moments(ctypes.byref(self.cfg), pos.ctypes.data, pos.shape[1], out.ctypes.data)
        return out


def moments_par(model, pos, inside_only=False):
    """(sum_w, sum_f [K]) of the Legendre basis over the walkers with all OpenMP threads; ``inside_only`` restricts the
    averages to the walkers inside the basis box (the estimator of ``VESBias_Ensemble(sample_domain='box')``)."""
    c = model.cfg
    K = (c.deg_x + 1) * ((c.deg_y + 1) if c.bias_kind == 2 else 1)
    out = np.zeros(K)
    sw = np.zeros(1)
    rc = load().oracle_moments_par(ctypes.byref(c), pos.ctypes.data, pos.shape[1], int(bool(inside_only)), sw.ctypes.data, out.ctypes.data)
    assert rc == 0
    return float(sw[0]), out


def num_threads():
    return int(load().oracle_num_threads())


def use_all_threads():
    """Use every hardware thread this process may run on (launchers such as torchrun set OMP_NUM_THREADS=1)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    load().oracle_set_num_threads(n)
    return num_threads()
