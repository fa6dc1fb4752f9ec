"""ctypes binding of the C ABI in ``include/pyves_b200.h`` (``csrc/libpyves_b200.so``).

There is no CPU fallback: if the shared library is missing, ``load()`` raises, and every compute
entry point needs a CUDA device.  ``Ensemble`` is a thin object wrapper around one handle.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libpyves_b200.so")

F32, F64 = 0, 1
SCHEME_OPENMM_LANGEVIN, SCHEME_MIDDLE = 0, 1
SCHEMES = {"openmm_langevin": SCHEME_OPENMM_LANGEVIN, "middle": SCHEME_MIDDLE}
(POT_SINGLE_WELL_1D, POT_DOUBLE_WELL_1D, POT_SINGLE_WELL_2D, POT_DOUBLE_WELL_2D, POT_SLIP_BOND_2D,
 POT_CATCH_BOND_2D, POT_SZABO_BEREZHKOVSKII, POT_MULLER_BROWN) = range(8)
BIAS_NONE, BIAS_LEGENDRE_1D, BIAS_LEGENDRE_2D, BIAS_MLP_1D, BIAS_PATHCV, BIAS_HARMONIC = range(6)
ACT_RELU, ACT_TANH = 0, 1
MAX_DEGREE = 63

OPT_SGD, OPT_RMSPROP, OPT_ADAM = 0, 1, 2
OPTIMIZERS = {"SGD": OPT_SGD, "RMSprop": OPT_RMSPROP, "Adam": OPT_ADAM}
GRAD_AVERAGE = {None: 0, "running": 1, "ewma": 2}

EINVAL, ECUDA, ESTATE, ENOMEM, EUNSUPPORTED = -1, -2, -3, -4, -5

_c = ctypes
_vp, _i, _i64, _u64, _d = _c.c_void_p, _c.c_int, _c.c_int64, _c.c_uint64, _c.c_double
_dp = _c.POINTER(_c.c_double)
_ip = _c.POINTER(_c.c_int)

# name -> (restype, argtypes); every symbol include/pyves_b200.h declares
SIGNATURES = {
    "pyves_create": (_i, [_c.POINTER(_vp), _i, _i64, _i, _u64, _i]),
    "pyves_destroy": (_i, [_vp]),
    "pyves_last_error": (_c.c_char_p, [_vp]),
    "pyves_abi_version": (_i, []),
    "pyves_set_integrator": (_i, [_vp, _d, _d, _d, _d, _i]),
    "pyves_set_potential": (_i, [_vp, _i, _dp, _i]),
    "pyves_set_walker_offset": (_i, [_vp, _i64]),
    "pyves_set_step_counter": (_i, [_vp, _i64]),
    "pyves_get_step_counter": (_i, [_vp, _c.POINTER(_i64)]),
    "pyves_set_bias_none": (_i, [_vp]),
    "pyves_set_bias_legendre1d": (_i, [_vp, _i, _i, _d, _d, _dp]),
    "pyves_set_bias_legendre2d": (_i, [_vp, _i, _i, _dp, _dp, _vp]),
    "pyves_set_bias_mlp": (_i, [_vp, _i, _d, _d, _i, _ip, _i, _dp, _vp]),
    "pyves_set_bias_pathcv": (_i, [_vp, _i, _i, _dp, _dp, _d, _dp, _d, _vp]),
    "pyves_set_bias_harmonic": (_i, [_vp, _i, _d, _d]),
    "pyves_basis_size": (_i, [_vp, _c.POINTER(_i64)]),
    "pyves_set_state": (_i, [_vp, _vp, _vp, _i, _vp]),
    "pyves_get_state": (_i, [_vp, _vp, _vp, _i, _vp]),
    "pyves_init_positions_mc": (_i, [_vp, _dp, _d, _i, _c.POINTER(_i64), _vp]),
    "pyves_init_velocities": (_i, [_vp, _vp]),
    "pyves_step": (_i, [_vp, _i64, _vp, _vp]),
    "pyves_step_record": (_i, [_vp, _i64, _i64, _i64, _vp, _vp, _vp]),
    "pyves_eval_bias": (_i, [_vp, _vp, _i64, _i, _vp, _vp, _i, _vp]),
    "pyves_energies": (_i, [_vp, _vp, _vp, _i, _vp]),
    "pyves_moments": (_i, [_vp, _vp, _i64, _i, _vp, _vp, _vp, _vp, _i, _vp]),
    "pyves_histogram": (_i, [_vp, _vp, _i64, _i, _i, _i, _dp, _ip, _vp, _vp, _i, _vp]),
    "pyves_set_tuning": (_i, [_vp, _i]),
    "pyves_step_wave_walkers": (_i, [_vp, _c.POINTER(_i64)]),
    "pyves_iterate_pinned": (_i, [_vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp]),
    "pyves_set_moment_domain": (_i, [_vp, _i]),
    "pyves_averaged_sgd_step": (_i, [_i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _d, _d, _i, _d, _i64, _vp, _vp, _vp]),
    "pyves_set_bias_legendre2d_device": (_i, [_vp, _i, _i, _dp, _vp, _vp]),
    "pyves_set_bias_legendre1d_device": (_i, [_vp, _i, _i, _d, _d, _vp, _vp]),
    "pyves_set_bias_like": (_i, [_vp, _vp, _vp]),
    "pyves_set_bias_mlp_device": (_i, [_vp, _i, _d, _d, _i, _ip, _i, _vp, _vp]),
    "pyves_optimizer_step": (_i, [_i, _i, _i64, _vp, _vp, _vp, _i, _d, _i64, _vp, _vp, _vp, _vp, _dp, _i64, _vp]),
    "pyves_comm_unique_id": (_i, [_vp]),
    "pyves_comm_create": (_i, [_c.POINTER(_vp), _i, _i, _i, _vp]),
    "pyves_comm_destroy": (_i, [_vp]),
    "pyves_allreduce_moments": (_i, [_vp, _vp, _i64, _vp]),
    "pyves_comm_nccl_version": (_i, [_c.POINTER(_i)]),
    "pyves_comm_last_error": (_c.c_char_p, []),
    "pyves_fp32_fma_peak": (_i, [_i, _dp]),
    "pyves_fp32_fma2_peak": (_i, [_i, _dp]),
    "pyves_launch_count": (_i64, []),
}

_lib = None


def load():
    """Load the shared library (once).  Raises if it has not been built -- there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "pyves_b200: %s is missing; build it with `python pyves_b200/csrc/build.py` "
                "(nvcc, sm_100a).  There is no CPU fallback." % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def _darr(a):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
    return a, a.ctypes.data_as(_dp)


def _raise(code, handle):
    msg = load().pyves_last_error(handle)
    msg = msg.decode() if msg else "pyves_b200 error %d" % code
    if code == EINVAL:
        raise ValueError(msg)
    if code == EUNSUPPORTED:
        raise NotImplementedError(msg)
    if code == ENOMEM:
        raise MemoryError(msg)
    raise RuntimeError(msg)


def averaged_sgd_step(device, sum_w, sum_wf, sum_wff, f_target, alpha, abar, scale, mu, beta, averaging, decay, n_after,
                      alpha_out, abar_out):
    """pyves_averaged_sgd_step on CUDA float64 tensors (enqueued on the current stream of ``device``)."""
    lib = load()
    rc = lib.pyves_averaged_sgd_step(device, alpha.numel(), _ptr(sum_w), _ptr(sum_wf), _ptr(sum_wff), _ptr(f_target), _ptr(alpha), _ptr(abar),
                                     _ptr(scale), float(mu), float(beta), 0 if averaging == 'ewma' else 1, float(decay), int(n_after),
                                     _ptr(alpha_out), _ptr(abar_out), _stream(device))
    if rc:
        raise (ValueError if rc == -1 else RuntimeError)(lib.pyves_last_error(None).decode())


def optimizer_step(device, kind, sum_w, sum_wf, f_target, grad_average, grad_decay, n_average, grad_state, theta, state1, state2, hyper, step):
    """pyves_optimizer_step on CUDA float64 tensors (in place on ``theta`` and the state vectors; enqueued on the current
    stream of ``device``).  ``kind``: 'SGD' | 'RMSprop' | 'Adam'; ``hyper``: (lr, a, b, eps) as in include/pyves_b200.h."""
    lib = load()
    hy = (_d * 4)(*[float(v) for v in hyper])
    rc = lib.pyves_optimizer_step(device, OPTIMIZERS[kind], theta.numel(), _ptr(sum_w), _ptr(sum_wf), _ptr(f_target), GRAD_AVERAGE[grad_average],
                                  float(grad_decay), int(n_average), _ptr(grad_state), _ptr(theta), _ptr(state1), _ptr(state2), hy, int(step),
                                  _stream(device))
    if rc:
        raise (ValueError if rc == -1 else RuntimeError)(lib.pyves_last_error(None).decode())


class Communicator:
    """The path's one collective through the C ABI (``pyves_comm_*`` / ``pyves_allreduce_moments``): an NCCL
    communicator over the ranks that share one ensemble.  ``from_process_group`` bootstraps it from an initialised
    ``torch.distributed`` group (rank 0's unique id is broadcast through it); any other channel that can ship 128 bytes
    works the same way (``unique_id()`` on rank 0, ``Communicator(device, nranks, rank, id)`` everywhere)."""

    def __init__(self, device, nranks, rank, unique_id):
        self.lib = load()
        self.device, self.nranks, self.rank = int(device), int(nranks), int(rank)
        buf = (ctypes.c_char * 128).from_buffer_copy(bytes(unique_id))
        self.h = _vp()
        rc = self.lib.pyves_comm_create(ctypes.byref(self.h), self.device, self.nranks, self.rank, ctypes.cast(buf, _vp))
        if rc:
            self.h = _vp()
            self._raise(rc)

    def _raise(self, rc):
        msg = self.lib.pyves_comm_last_error().decode()
        raise (ValueError if rc == EINVAL else NotImplementedError if rc == EUNSUPPORTED else RuntimeError)(msg)

    @staticmethod
    def unique_id():
        lib = load()
        buf = (ctypes.c_char * 128)()
        rc = lib.pyves_comm_unique_id(ctypes.cast(buf, _vp))
        if rc:
            raise RuntimeError(lib.pyves_comm_last_error().decode())
        return bytes(buf)

    @staticmethod
    def nccl_version():
        v = _i(0)
        rc = load().pyves_comm_nccl_version(ctypes.byref(v))
        return v.value if rc == 0 else None

    @classmethod
    def from_process_group(cls, device, group=None):
        import torch
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        box = [cls.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        torch.cuda.synchronize(device)
        return cls(device, world, rank, box[0])

    def allreduce_(self, flat):
        """In-place sum of a contiguous CUDA float64 tensor over the ranks, enqueued on the current stream."""
        import torch
        if not (torch.is_tensor(flat) and flat.is_cuda and flat.dtype == torch.float64 and flat.is_contiguous() and flat.device.index == self.device):
            raise ValueError("allreduce_ needs a contiguous CUDA float64 tensor on device %d" % self.device)
        rc = self.lib.pyves_allreduce_moments(self.h, ctypes.c_void_p(flat.data_ptr()), flat.numel(), _stream(self.device))
        if rc:
            self._raise(rc)
        return flat

    def close(self):
        if getattr(self, "h", None):
            self.lib.pyves_comm_destroy(self.h)
            self.h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def launch_count():
    return int(load().pyves_launch_count())


def fp32_fma_peak(device=0, packed=False):
    out = _d(0.0)
    fn = load().pyves_fp32_fma2_peak if packed else load().pyves_fp32_fma_peak
    rc = fn(device, ctypes.byref(out))
    if rc:
        _raise(rc, None)
    return out.value


def _ptr(t):
    """Raw address of a torch tensor / numpy array / None / int."""
    if t is None:
        return None
    if isinstance(t, int):
        return t
    if isinstance(t, np.ndarray):
        return t.ctypes.data
    return t.data_ptr()


def _stream(device):
    import torch
    return torch.cuda.current_stream(device).cuda_stream


class _Range:
    """NVTX range around a library call when PYVES_NVTX=1 (ncu --nvtx / nsys filter on pyves.step, pyves.moments,
    pyves.update); a no-op otherwise."""
    enabled = os.environ.get("PYVES_NVTX", "0") == "1"

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if _Range.enabled:
            import torch
            torch.cuda.nvtx.range_push(self.name)

    def __exit__(self, *exc):
        if _Range.enabled:
            import torch
            torch.cuda.nvtx.range_pop()
        return False


class Ensemble:
    """One ensemble of walkers on one GPU (one ``pyves_t`` handle)."""

    def __init__(self, n_walkers, dims=2, seed=0, precision="fp32", device=0):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("pyves_b200 needs a CUDA device; there is no CPU fallback")
        self.lib = load()
        self.n = int(n_walkers)
        self.dims = int(dims)
        self.precision = {"fp32": F32, "fp64": F64, F32: F32, F64: F64}[precision]
        self.device = int(device)
        self.torch_dtype = torch.float64 if self.precision == F64 else torch.float32
        self.np_dtype = np.float64 if self.precision == F64 else np.float32
        self.h = _vp()
        rc = self.lib.pyves_create(ctypes.byref(self.h), self.device, self.n, self.dims, int(seed) & (2 ** 64 - 1),
                                   self.precision)
        if rc:
            self.h = _vp()
            _raise(rc, None)

    # -------------------------------------------------------------------------------------- plumbing
    def _ck(self, rc):
        if rc:
            _raise(rc, self.h)

    def close(self):
        if getattr(self, "h", None):
            self.lib.pyves_destroy(self.h)
            self.h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def stream(self):
        return _stream(self.device)

    def _check(self, t, shape=None, dtype=None):
        """A caller tensor must be contiguous, on this device, of the handle's dtype."""
        import torch
        if not isinstance(t, torch.Tensor) or not t.is_cuda or t.device.index != self.device:
            raise ValueError("expected a CUDA tensor on device %d" % self.device)
        if not t.is_contiguous():
            raise ValueError("expected a contiguous tensor")
        if t.dtype != (dtype or self.torch_dtype):
            raise ValueError("expected dtype %s" % (dtype or self.torch_dtype))
        if shape is not None and tuple(t.shape) != tuple(shape):
            raise ValueError("expected shape %s, got %s" % (tuple(shape), tuple(t.shape)))
        return t

    # ---------------------------------------------------------------------------------- configuration
    def set_integrator(self, mass, kT, friction, dt, scheme="openmm_langevin"):
        self._ck(self.lib.pyves_set_integrator(self.h, mass, kT, friction, dt, SCHEMES.get(scheme, scheme)))

    def set_potential(self, kind, params=()):
        a, p = _darr(params)
        self._ck(self.lib.pyves_set_potential(self.h, kind, p if a.size else None, a.size))

    def set_walker_offset(self, gid0):
        self._ck(self.lib.pyves_set_walker_offset(self.h, gid0))

    @property
    def step_counter(self):
        out = _i64(0)
        self._ck(self.lib.pyves_get_step_counter(self.h, ctypes.byref(out)))
        return out.value

    @step_counter.setter
    def step_counter(self, v):
        self._ck(self.lib.pyves_set_step_counter(self.h, v))

    def set_tuning(self, variant):
        self._ck(self.lib.pyves_set_tuning(self.h, variant))

    def iterate_pinned(self, pos, vel, nsteps, out, stream=None):
        """One iteration with HOST buffers, everything enqueued on the stream (no synchronisation): the walker state from the PINNED
        numpy arrays ``pos`` / ``vel`` ([dims, n], the handle's dtype, C-contiguous), ``nsteps`` steps, the basis moments into the CUDA
        tensors ``out = (sum_w, sum_wf)``, the walker state back into ``pos`` / ``vel``.  One C call instead of set_state(async) /
        step / moments / get_state(async); synchronise the stream before reading the arrays."""
        for a in (pos, vel):
            if not (isinstance(a, np.ndarray) and a.dtype == self.np_dtype and a.flags["C_CONTIGUOUS"] and a.shape == (self.dims, self.n)):
                raise ValueError("iterate_pinned needs C-contiguous numpy arrays [dims, n] of the handle's dtype")
        st = self.stream if stream is None else stream
        self._ck(self.lib.pyves_iterate_pinned(self.h, pos.ctypes.data, vel.ctypes.data, int(nsteps), out[0].data_ptr(), out[1].data_ptr(),
                                               pos.ctypes.data, vel.ctypes.data, st))

    def wave_walkers(self):
        """Walkers one full wave of this handle's step kernel advances (0: no granularity worth respecting).  Chunked
        pipelines should cut the ensemble at multiples of it: a partial wave costs a whole one."""
        n = _i64(0)
        self._ck(self.lib.pyves_step_wave_walkers(self.h, _c.byref(n)))
        return int(n.value)

    def set_moment_domain(self, inside_only):
        """True: ``moments()`` averages over the rows inside the basis interval only (conditional
        ensemble average on the support of the target); False (default): every row, clamped."""
        self._ck(self.lib.pyves_set_moment_domain(self.h, int(bool(inside_only))))

    def set_bias_none(self):
        self._ck(self.lib.pyves_set_bias_none(self.h))

    def set_bias_legendre1d(self, axis, degree, lo, hi, w):
        a, p = _darr(w)
        if a.size != degree + 1:
            raise ValueError("weights must have degree + 1 entries")
        self._ck(self.lib.pyves_set_bias_legendre1d(self.h, axis, degree, lo, hi, p))

    def set_bias_legendre2d(self, deg_x, deg_y, minmax, w):
        a, p = _darr(w)
        if a.size != (deg_x + 1) * (deg_y + 1):
            raise ValueError("weights must have (deg_x + 1) * (deg_y + 1) entries")
        m, mp = _darr(minmax)
        self._ck(self.lib.pyves_set_bias_legendre2d(self.h, deg_x, deg_y, mp, p, self.stream))

    def set_bias_legendre1d_device(self, axis, degree, lo, hi, w_dev):
        """1D coefficients from a CUDA float64 tensor [degree + 1], read in stream order."""
        import torch
        if not (torch.is_tensor(w_dev) and w_dev.is_cuda and w_dev.dtype == torch.float64 and w_dev.is_contiguous()):
            raise ValueError("w_dev must be a contiguous CUDA float64 tensor")
        if w_dev.numel() != degree + 1:
            raise ValueError("weights must have degree + 1 entries")
        if w_dev.device.index != self.device:
            raise ValueError("w_dev lives on another device")
        self._ck(self.lib.pyves_set_bias_legendre1d_device(self.h, axis, degree, lo, hi, ctypes.c_void_p(w_dev.data_ptr()), self.stream))

    def set_bias_legendre2d_device(self, deg_x, deg_y, minmax, w_dev):
        """Coefficients from a CUDA float64 tensor [deg_x + 1, deg_y + 1] (or flat), read in stream order:
        no host round trip between the optimiser and the next launch."""
        import torch
        if not (torch.is_tensor(w_dev) and w_dev.is_cuda and w_dev.dtype == torch.float64 and w_dev.is_contiguous()):
            raise ValueError("w_dev must be a contiguous CUDA float64 tensor")
        if w_dev.numel() != (deg_x + 1) * (deg_y + 1):
            raise ValueError("weights must have (deg_x + 1) * (deg_y + 1) entries")
        if w_dev.device.index != self.device:
            raise ValueError("w_dev lives on another device")
        m, mp = _darr(minmax)
        self._ck(self.lib.pyves_set_bias_legendre2d_device(self.h, deg_x, deg_y, mp, ctypes.c_void_p(w_dev.data_ptr()), self.stream))

    def set_bias_like(self, leader):
        """Share the device-resident Legendre coefficients of ``leader`` (same device and precision)."""
        self._ck(self.lib.pyves_set_bias_like(self.h, leader.h, self.stream))

    def set_bias_mlp(self, axis, lo, hi, sizes, act, theta):
        a, p = _darr(theta)
        s = np.ascontiguousarray(np.asarray(sizes, dtype=np.int32))
        dims = [1] + [int(v) for v in s] + [1]
        if len(s) and a.size != sum(dims[i] * dims[i + 1] + dims[i + 1] for i in range(len(dims) - 1)):
            raise ValueError("parameter vector does not match the layer sizes")
        self._ck(self.lib.pyves_set_bias_mlp(self.h, axis, lo, hi, len(s), s.ctypes.data_as(_ip), act, p, self.stream))

    def set_bias_mlp_device(self, axis, lo, hi, sizes, act, theta_dev):
        """NN parameters from a CUDA float64 tensor (torch ``parameters()`` order), read in stream order; the fold of the
        first two layers and the repack for the step kernels run on the device."""
        import torch
        if not (torch.is_tensor(theta_dev) and theta_dev.is_cuda and theta_dev.dtype == torch.float64 and theta_dev.is_contiguous()):
            raise ValueError("theta_dev must be a contiguous CUDA float64 tensor")
        if theta_dev.device.index != self.device:
            raise ValueError("theta_dev lives on another device")
        s = np.ascontiguousarray(np.asarray(sizes, dtype=np.int32))
        dims = [1] + [int(v) for v in s] + [1]
        if theta_dev.numel() != sum(dims[i] * dims[i + 1] + dims[i + 1] for i in range(len(dims) - 1)):
            raise ValueError("parameter vector does not match the layer sizes")
        self._ck(self.lib.pyves_set_bias_mlp_device(self.h, axis, lo, hi, len(s), s.ctypes.data_as(_ip), act, ctypes.c_void_p(theta_dev.data_ptr()),
                                                    self.stream))

    def set_bias_pathcv(self, degree, x_i, y_i, lam, w, phi=0.0):
        xa, xp = _darr(x_i)
        ya, yp = _darr(y_i)
        wa, wp = _darr(w)
        if xa.size != ya.size:
            raise ValueError("x_i and y_i must have the same length")
        if wa.size != degree + 1:
            raise ValueError("weights must have degree + 1 entries")
        self._ck(self.lib.pyves_set_bias_pathcv(self.h, degree, xa.size, xp, yp, lam, wp, phi, self.stream))

    def set_bias_harmonic(self, axis, k, s0):
        self._ck(self.lib.pyves_set_bias_harmonic(self.h, axis, k, s0))

    @property
    def basis_size(self):
        out = _i64(0)
        self._ck(self.lib.pyves_basis_size(self.h, ctypes.byref(out)))
        return out.value

    # ------------------------------------------------------------------------------------------ state
    def set_state(self, pos=None, vel=None, async_copy=False):
        """SoA arrays [dims, n]: CUDA tensors (device copy) or numpy arrays (host copy).  ``async_copy``: the numpy
        arrays are PINNED, the copy is only enqueued on the current stream (synchronise before touching them)."""
        host = isinstance(pos if pos is not None else vel, np.ndarray)
        keep = []
        for t in (pos, vel):
            if t is None:
                keep.append(None)
            elif host:
                keep.append(np.ascontiguousarray(t, dtype=self.np_dtype).reshape(self.dims, self.n))
            else:
                keep.append(self._check(t, (self.dims, self.n)))
        self._ck(self.lib.pyves_set_state(self.h, _ptr(keep[0]), _ptr(keep[1]), (2 if async_copy else 1) if host else 0, self.stream))

    def get_state(self, host=False, out=None, async_copy=False):
        """(pos, vel) as CUDA tensors, numpy arrays (``host``) or into ``out``; ``async_copy`` as in ``set_state``."""
        import torch
        if out is not None:
            pos, vel = out
        elif host:
            pos = np.empty((self.dims, self.n), dtype=self.np_dtype)
            vel = np.empty((self.dims, self.n), dtype=self.np_dtype)
        else:
            pos = torch.empty((self.dims, self.n), dtype=self.torch_dtype, device="cuda:%d" % self.device)
            vel = torch.empty_like(pos)
        host = isinstance(pos, np.ndarray)
        self._ck(self.lib.pyves_get_state(self.h, _ptr(pos), _ptr(vel), (2 if async_copy else 1) if host else 0, self.stream))
        return pos, vel

    def init_positions_mc(self, box, beta, ntrials=100):
        b, bp = _darr(box)
        nf = _i64(0)
        self._ck(self.lib.pyves_init_positions_mc(self.h, bp, beta, ntrials, ctypes.byref(nf), self.stream))
        return nf.value

    def init_velocities(self):
        self._ck(self.lib.pyves_init_velocities(self.h, self.stream))

    # --------------------------------------------------------------------------------------- hot path
    def step(self, nsteps, noise=None):
        if noise is not None:
            self._check(noise, (nsteps, self.dims, self.n))
        with _Range("pyves.step"):
            self._ck(self.lib.pyves_step(self.h, nsteps, _ptr(noise), self.stream))

    def step_record(self, nsteps, every=1, n_rec=1, noise=None):
        import torch
        if noise is not None:
            self._check(noise, (nsteps, self.dims, self.n))
        nframes = (nsteps + every - 1) // every
        traj = torch.empty((nframes, self.dims, n_rec), dtype=self.torch_dtype, device="cuda:%d" % self.device)
        self._ck(self.lib.pyves_step_record(self.h, nsteps, every, n_rec, _ptr(traj), _ptr(noise), self.stream))
        return traj

    def eval_bias(self, pos, want_energy=True, want_force=True):
        """pos [n, ncols] -> (E [n], F [n, ncols]); torch CUDA tensors or numpy (host staging)."""
        import torch
        host = isinstance(pos, np.ndarray)
        if host:
            pos = np.ascontiguousarray(pos, dtype=self.np_dtype)
            n, ncols = pos.shape
            E = np.empty(n, dtype=self.np_dtype) if want_energy else None
            F = np.empty((n, ncols), dtype=self.np_dtype) if want_force else None
        else:
            self._check(pos)
            n, ncols = pos.shape
            E = torch.empty(n, dtype=self.torch_dtype, device=pos.device) if want_energy else None
            F = torch.empty((n, ncols), dtype=self.torch_dtype, device=pos.device) if want_force else None
        self._ck(self.lib.pyves_eval_bias(self.h, _ptr(pos), n, ncols, _ptr(E), _ptr(F), int(host), self.stream))
        return E, F

    def energies(self, host=False):
        import torch
        if host:
            PE = np.empty(self.n, dtype=self.np_dtype)
            KE = np.empty(self.n, dtype=self.np_dtype)
        else:
            PE = torch.empty(self.n, dtype=self.torch_dtype, device="cuda:%d" % self.device)
            KE = torch.empty_like(PE)
        self._ck(self.lib.pyves_energies(self.h, _ptr(PE), _ptr(KE), int(host), self.stream))
        return PE, KE

    def moments(self, pos=None, row_weights=None, cov=False, host=False, out=None):
        """(sum_w, sum_wf [K], sum_wff [K, K] or None) of the walkers (pos None) or of rows pos [n, ncols].
        ``out`` = (sum_w, sum_wf, sum_wff or None): caller-owned CUDA float64 tensors that are overwritten (no
        allocation, e.g. views of one flat buffer that is then all-reduced in place)."""
        import torch
        if out is not None:
            sw, sf, sff = out
            n, ncols = (0, 0) if pos is None else pos.shape
            with _Range("pyves.moments"):
                self._ck(self.lib.pyves_moments(self.h, _ptr(pos), n, ncols, _ptr(row_weights), _ptr(sw), _ptr(sf), _ptr(sff), 0, self.stream))
            return sw, sf, sff
        K = self.basis_size
        n, ncols = (0, 0) if pos is None else pos.shape
        if pos is not None and isinstance(pos, np.ndarray):
            host = True
            pos = np.ascontiguousarray(pos, dtype=self.np_dtype)
        if row_weights is not None and isinstance(row_weights, np.ndarray):
            host = True
            row_weights = np.ascontiguousarray(row_weights, dtype=self.np_dtype)
        if host:
            sw = np.zeros(1)
            sf = np.zeros(K)
            sff = np.zeros((K, K)) if cov else None
        else:
            dev = "cuda:%d" % self.device
            if pos is not None:
                self._check(pos)
            if row_weights is not None:
                self._check(row_weights)
            sw = torch.zeros(1, dtype=torch.float64, device=dev)
            sf = torch.zeros(K, dtype=torch.float64, device=dev)
            sff = torch.zeros((K, K), dtype=torch.float64, device=dev) if cov else None
        if pos is not None and n == 0:
            return sw, sf, sff           # an empty row set sums to zero (a NULL pointer would mean "the handle's walkers")
        with _Range("pyves.moments"):
            self._ck(self.lib.pyves_moments(self.h, _ptr(pos), n, ncols, _ptr(row_weights), _ptr(sw), _ptr(sf), _ptr(sff),
                                            int(host), self.stream))
        return sw, sf, sff

    def histogram(self, ranges, nbins, axis=0, pos=None, row_weights=None, counts=None, host=False):
        """1D (nbins int) or 2D (nbins pair) histogram, accumulated into ``counts`` (float64)."""
        import torch
        ndim = 1 if np.isscalar(nbins) else 2
        nb = np.ascontiguousarray(np.atleast_1d(nbins).astype(np.int32))
        r, rp = _darr(np.ravel(ranges))
        shape = (int(nb[0]),) if ndim == 1 else (int(nb[0]), int(nb[1]))
        n, ncols = (0, 0) if pos is None else pos.shape
        if (pos is not None and isinstance(pos, np.ndarray)) or isinstance(counts, np.ndarray):
            host = True
        if counts is None:
            counts = np.zeros(shape) if host else torch.zeros(shape, dtype=torch.float64, device="cuda:%d" % self.device)
        if host and pos is not None:
            pos = np.ascontiguousarray(pos, dtype=self.np_dtype)
        if host and row_weights is not None:
            row_weights = np.ascontiguousarray(row_weights, dtype=self.np_dtype)
        if pos is not None and n == 0:
            return counts                # nothing to add (a NULL pointer would mean "the handle's walkers")
        self._ck(self.lib.pyves_histogram(self.h, _ptr(pos), n, ncols, ndim, axis, rp, nb.ctypes.data_as(_ip),
                                          _ptr(row_weights), _ptr(counts), int(host), self.stream))
        return counts
