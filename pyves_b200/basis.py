"""Basis expansions -- ``torch.nn.Module`` mirror of ``ves/basis.py`` evaluated by the CUDA library.

Same classes, constructor keywords, parameter names and error messages as the reference:
``LegendreBasis1D`` (``ves/basis.py:16-156``), ``NNBasis1D`` (``:159-232``),
``LegendreBasis2DPathCV`` (``:386-548``), plus the 2D tensor-product ``LegendreBasis2D`` that
``BASELINE.json`` asks for.  ``forward(positions) -> energy`` accepts the shapes the reference's
examples use (``[N, 3]``, ``[M, 2]``, ``[M, 1]`` and, for the NN, ``[M, 1, 1]``), on CUDA or on the
host (host tensors go through the host-buffer entry points of the C ABI).  Energies, position
gradients (forces) and parameter gradients all come from the sm_100a kernels
(``pyves_eval_bias`` / ``pyves_moments``) through a ``torch.autograd.Function``; there is no
pure-torch compute path and a missing library / GPU raises.

``export_torchscript()`` returns a scripted, framework-free definition of the same function for the
``model_loc`` files the reference's tooling reloads (``ves/bias.py:52-53,249-253``); it is an export
format, never used for computation here.
"""
import numpy as np
import torch
import torch.nn as nn

from . import _lib

_AXES = {'x': 0, 'y': 1}


def _axis_id(axis):
    if axis not in _AXES:
        raise ValueError("Invalid axis")
    return _AXES[axis]


class _KernelBias(torch.autograd.Function):
    """energy = V(positions; params) with dV/dpositions and dV/dparams from the CUDA kernels."""

    @staticmethod
    def forward(ctx, module, pos, *params):
        ens, host = module._evaluator(pos)
        module._upload(ens, params)
        if host:
            E, F = ens.eval_bias(pos.detach().numpy())
            E, F = torch.from_numpy(E), torch.from_numpy(F)
        else:
            E, F = ens.eval_bias(pos.detach())
        ctx.module = module
        ctx.host = host
        ctx.save_for_backward(pos.detach(), F, *[p.detach() for p in params])
        return E

    @staticmethod
    def backward(ctx, g):
        pos, F, *params = ctx.saved_tensors
        module = ctx.module
        grad_pos = -(F * g.unsqueeze(-1)) if ctx.needs_input_grad[1] else None
        grad_params = [None] * len(params)
        if any(ctx.needs_input_grad[2:]):
            ens, _ = module._evaluator(pos)
            module._upload(ens, params)
            g = g.contiguous()
            if ctx.host:
                _, sf, _ = ens.moments(pos=pos.numpy(), row_weights=g.numpy())
                sf = torch.from_numpy(sf)
            else:
                _, sf, _ = ens.moments(pos=pos, row_weights=g)
            o = 0
            for i, p in enumerate(params):
                if ctx.needs_input_grad[2 + i]:
                    grad_params[i] = sf[o:o + p.numel()].reshape(p.shape).to(device=p.device, dtype=p.dtype)
                o += p.numel()
        return (None, grad_pos) + tuple(grad_params)


class _KernelModule(nn.Module):
    """Shared plumbing: one evaluator handle per (device, precision), shape handling."""
    min_cols = 1

    def _evaluator(self, pos):
        host = not pos.is_cuda
        dev = torch.cuda.current_device() if host else pos.device.index   # host rows: this process's GPU, not GPU 0
        prec = "fp32" if pos.dtype == torch.float32 else "fp64"
        cache = self.__dict__.setdefault("_evaluators", {})
        if (dev, prec) not in cache:
            cache[(dev, prec)] = _lib.Ensemble(1, dims=2, precision=prec, device=dev)
        return cache[(dev, prec)], host

    def _prepare(self, positions):
        if not torch.is_tensor(positions):
            positions = torch.as_tensor(positions)
        if positions.dtype not in (torch.float32, torch.float64):
            positions = positions.to(torch.float64)
        if positions.dim() != 2 or positions.shape[1] < self.min_cols:
            raise ValueError("positions must have shape [N, C] with C >= %d" % self.min_cols)
        if positions.shape[1] > 3:
            positions = positions[:, :3]
        return positions.contiguous()

    def _run(self, positions):
        return _KernelBias.apply(self, self._prepare(positions), *self._kernel_params())

    def __getstate__(self):
        d = self.__dict__.copy()
        d.pop("_evaluators", None)      # device handles are not picklable
        return d


# ================================================================================================
class LegendreBasis1D(_KernelModule):
    r"""Legendre expansion along x or y: $x' = (x - (min+max)/2) / ((max-min)/2)$ clamped to
    $[-1, 1]$, $B(x') = \sum_{i=0}^{d} w_i P_i(x')$ (``ves/basis.py:16-156``).

    Attributes: degree, min, max, axis ('x' | 'y'), weights (``nn.Parameter``, float64, len degree+1).
    """

    def __init__(self, degree, min, max, axis='x', weights=None):
        super().__init__()
        if degree > _lib.MAX_DEGREE:
            raise ValueError("degree must be <= %d" % _lib.MAX_DEGREE)
        self.degree = degree
        self.min = min
        self.max = max
        self.axis = axis
        if weights is not None:
            w = torch.from_numpy(np.asarray(weights)).to(torch.float64).clone()
        else:
            w = torch.rand(degree + 1, dtype=torch.float64)
        self.weights = nn.Parameter(w)

    @classmethod
    def legendre_polynomial_expansion(cls, x, degree: int, weights: torch.Tensor) -> torch.Tensor:
        """sum_n w_n P_n(x) for x already in [-1, 1] (``ves/basis.py:86-121``, Bonnet recursion on the GPU)."""
        m = LegendreBasis1D(degree, -1.0, 1.0, weights=np.zeros(degree + 1))
        out = _KernelBias.apply(m, m._prepare(x.reshape(-1, 1)), weights.to(torch.float64))
        return out.reshape(x.shape)

    @classmethod
    def legendre_polynomial(cls, x, degree: int) -> torch.Tensor:
        """Single P_degree(x) (deprecated upstream, ``ves/basis.py:56-84``)."""
        w = torch.zeros(degree + 1, dtype=torch.float64)
        w[degree] = 1.0
        return cls.legendre_polynomial_expansion(x, degree, w)

    def _kernel_params(self):
        return (self.weights,)

    def _upload(self, ens, params):
        ens.set_bias_legendre1d(_axis_id(self.axis), self.degree, float(self.min), float(self.max),
                                params[0].detach().cpu().numpy())

    def forward(self, positions):
        """positions [N, C] (column 0 = x, column 1 = y) -> energy [N] in kJ/mol."""
        a = _axis_id(self.axis)
        positions = self._prepare(positions)
        if positions.shape[1] <= a:
            raise ValueError("positions have no column for axis %r" % self.axis)
        return _KernelBias.apply(self, positions, self.weights)

    def descriptor(self):
        return ("legendre1d", dict(axis=_axis_id(self.axis), degree=self.degree, lo=float(self.min), hi=float(self.max),
                                   w=self.weights.detach().cpu().numpy().copy()))

    def export_torchscript(self):
        return torch.jit.script(_ExportLegendre1D(self.degree, float(self.min), float(self.max), _axis_id(self.axis),
                                                  self.weights.detach().cpu().clone()))


class LegendreBasis2D(_KernelModule):
    r"""Tensor-product Legendre expansion $V = \sum_{ij} w_{ij} P_i(x') P_j(y')$ with the per-axis
    scaling / clamp conventions of ``LegendreBasis1D`` (new; ``BASELINE.json`` configs[1]).

    weights: ``nn.Parameter`` [degree_x + 1, degree_y + 1], float64 (zeros by default: VES starts unbiased).
    """
    min_cols = 2

    def __init__(self, degree_x, degree_y, min_x, max_x, min_y, max_y, weights=None):
        super().__init__()
        if max(degree_x, degree_y) > _lib.MAX_DEGREE:
            raise ValueError("degree must be <= %d" % _lib.MAX_DEGREE)
        self.degree_x, self.degree_y = degree_x, degree_y
        self.min_x, self.max_x, self.min_y, self.max_y = min_x, max_x, min_y, max_y
        shape = (degree_x + 1, degree_y + 1)
        if weights is not None:
            w = torch.from_numpy(np.asarray(weights, dtype=np.float64).reshape(shape)).clone()
        else:
            w = torch.zeros(shape, dtype=torch.float64)
        self.weights = nn.Parameter(w)

    def _kernel_params(self):
        return (self.weights,)

    def _minmax(self):
        return (float(self.min_x), float(self.max_x), float(self.min_y), float(self.max_y))

    def _upload(self, ens, params):
        ens.set_bias_legendre2d(self.degree_x, self.degree_y, self._minmax(), params[0].detach().cpu().numpy())

    def forward(self, positions):
        return self._run(positions)

    def descriptor(self):
        return ("legendre2d", dict(deg_x=self.degree_x, deg_y=self.degree_y, minmax=self._minmax(),
                                   w=self.weights.detach().cpu().numpy().copy()))

    def export_torchscript(self):
        return torch.jit.script(_ExportLegendre2D(self.degree_x, self.degree_y, self._minmax(), self.weights.detach().cpu().clone()))


class NNBasis1D(_KernelModule):
    r"""Neural-network expansion along x or y: $x' = (x - min) / (max - min)$, $B(x') = N(x')$
    (``ves/basis.py:159-232``).  Layer structure as upstream: ``Linear(1, h0)`` without activation,
    then ``Linear(h_{l-1}, h_l)`` + act, then ``Linear(h_last, 1)``.
    """

    def __init__(self, min, max, axis='x', hidden_layer_sizes=[], act=nn.ReLU):
        super().__init__()
        self.min = min
        self.max = max
        self.axis = axis
        self.layers = nn.ModuleList()
        if len(hidden_layer_sizes) == 0:
            raise ValueError("NNBasis needs at least one hidden layer")
        if act is nn.ReLU:
            self._act = _lib.ACT_RELU
        elif act is nn.Tanh:
            self._act = _lib.ACT_TANH
        else:
            raise NotImplementedError("the CUDA path implements nn.ReLU and nn.Tanh activations")
        self.hidden_layer_sizes = [int(h) for h in hidden_layer_sizes]
        sizes = self.hidden_layer_sizes
        self.layers.append(nn.Linear(1, sizes[0], dtype=torch.float64))
        for lidx in range(1, len(sizes)):
            self.layers.append(nn.Linear(sizes[lidx - 1], sizes[lidx], dtype=torch.float64))
            self.layers.append(act())
        self.layers.append(nn.Linear(sizes[-1], 1, dtype=torch.float64))

    def _kernel_params(self):
        return tuple(self.parameters())

    def _theta(self, params):
        return np.concatenate([p.detach().cpu().numpy().ravel() for p in params])

    def _upload(self, ens, params):
        ens.set_bias_mlp(_axis_id(self.axis), float(self.min), float(self.max), self.hidden_layer_sizes, self._act,
                         self._theta(params))

    def forward(self, positions):
        """positions [N, C] -> energy [N]; the reference's ``[M, 1, 1]`` fitting input -> ``[M, 1]``."""
        a = _axis_id(self.axis)
        if torch.is_tensor(positions) and positions.dim() == 3:
            flat = positions.reshape(positions.shape[0], -1)
            return self.forward(flat).unsqueeze(-1)
        positions = self._prepare(positions)
        if positions.shape[1] <= a:
            raise ValueError("positions have no column for axis %r" % self.axis)
        return _KernelBias.apply(self, positions, *self._kernel_params())

    def descriptor(self):
        return ("mlp", dict(axis=_axis_id(self.axis), lo=float(self.min), hi=float(self.max), sizes=list(self.hidden_layer_sizes),
                            act=self._act, theta=self._theta(self._kernel_params())))

    def export_torchscript(self):
        return torch.jit.script(_ExportNN(float(self.min), float(self.max), _axis_id(self.axis), self.layers))


class LegendreBasis2DRadialCV(nn.Module):
    """Deprecated upstream in favour of the path CV (``ves/basis.py:239``); not provided."""

    def __init__(self, *args, **kwargs):
        raise NotImplementedError("LegendreBasis2DRadialCV is deprecated upstream; use LegendreBasis2DPathCV instead.")


class LegendreBasis2DPathCV(_KernelModule):
    r"""Legendre expansion along the path CV $s$ of the images $(x_i, y_i)$ plus a linear restraint
    $\phi z$ on the distance from the path (``ves/basis.py:386-548``):
    $s = \frac{1}{N} \frac{\sum_i (i+1) e^{-\lambda d_i}}{\sum_i e^{-\lambda d_i}}$,
    $z = -\frac{1}{\lambda} \ln \sum_i e^{-\lambda d_i}$, $s' = 2s - 1$ clamped to $[-1, 1]$.
    """
    min_cols = 2

    def __init__(self, degree, x_i, y_i, lam, weights=None, phi=0):
        super().__init__()
        if degree > _lib.MAX_DEGREE:
            raise ValueError("degree must be <= %d" % _lib.MAX_DEGREE)
        self.degree = degree
        assert len(x_i) == len(y_i)
        self.Npath = len(x_i)
        self.x_i = torch.from_numpy(np.asarray(x_i, dtype=np.float64)).clone()
        self.y_i = torch.from_numpy(np.asarray(y_i, dtype=np.float64)).clone()
        self.lam = lam
        if weights is not None:
            w = torch.from_numpy(np.asarray(weights)).to(torch.float64).clone()
        else:
            w = torch.rand(degree + 1, dtype=torch.float64)
        self.weights = nn.Parameter(w)
        self.phi = phi

    # the reference class carries its own copies of the two Legendre helpers (``ves/basis.py:442-507``)
    legendre_polynomial_expansion = LegendreBasis1D.__dict__['legendre_polynomial_expansion']
    legendre_polynomial = LegendreBasis1D.__dict__['legendre_polynomial']

    def _kernel_params(self):
        return (self.weights,)

    def _upload(self, ens, params):
        ens.set_bias_pathcv(self.degree, self.x_i.numpy(), self.y_i.numpy(), float(self.lam),
                            params[0].detach().cpu().numpy(), float(self.phi))

    def forward(self, positions):
        return self._run(positions)

    def descriptor(self):
        return ("pathcv", dict(degree=self.degree, x_i=self.x_i.numpy().copy(), y_i=self.y_i.numpy().copy(), lam=float(self.lam),
                               w=self.weights.detach().cpu().numpy().copy(), phi=float(self.phi)))

    def export_torchscript(self):
        return torch.jit.script(_ExportPathCV(self.degree, self.x_i.clone(), self.y_i.clone(), float(self.lam),
                                              self.weights.detach().cpu().clone(), float(self.phi)))


# ================================================================================================
# TorchScript export definitions (file format for model_loc; not a compute path of this package)
def _legendre_stack(s: torch.Tensor, degree: int) -> torch.Tensor:
    """[degree + 1, N] table of P_n(s) by Bonnet's recursion."""
    cols = [torch.ones_like(s)]
    if degree >= 1:
        cols.append(s)
    for n in range(1, degree):
        cols.append(((2 * n + 1) * s * cols[n] - n * cols[n - 1]) / (n + 1))
    return torch.stack(cols, 0)


class _ExportLegendre1D(nn.Module):
    def __init__(self, degree: int, lo: float, hi: float, axis: int, weights):
        super().__init__()
        self.degree, self.lo, self.hi, self.axis = degree, lo, hi, axis
        self.weights = nn.Parameter(weights)

    def forward(self, positions):
        s = (positions[:, self.axis] - 0.5 * (self.lo + self.hi)) / (0.5 * (self.hi - self.lo))
        table = _legendre_stack(torch.clamp(s, -1.0, 1.0), self.degree)
        return (self.weights.to(table.dtype).unsqueeze(1) * table).sum(0)


class _ExportLegendre2D(nn.Module):
    def __init__(self, degree_x: int, degree_y: int, minmax, weights):
        super().__init__()
        self.degree_x, self.degree_y = degree_x, degree_y
        self.xlo, self.xhi, self.ylo, self.yhi = minmax
        self.weights = nn.Parameter(weights)

    def forward(self, positions):
        sx = torch.clamp((positions[:, 0] - 0.5 * (self.xlo + self.xhi)) / (0.5 * (self.xhi - self.xlo)), -1.0, 1.0)
        sy = torch.clamp((positions[:, 1] - 0.5 * (self.ylo + self.yhi)) / (0.5 * (self.yhi - self.ylo)), -1.0, 1.0)
        tx = _legendre_stack(sx, self.degree_x)
        ty = _legendre_stack(sy, self.degree_y)
        return (tx * (self.weights.to(tx.dtype) @ ty)).sum(0)


class _ExportNN(nn.Module):
    def __init__(self, lo: float, hi: float, axis: int, layers):
        super().__init__()
        self.lo, self.hi, self.axis = lo, hi, axis
        self.layers = layers

    def forward(self, positions):
        h = ((positions[:, self.axis] - self.lo) / (self.hi - self.lo)).unsqueeze(-1)
        for layer in self.layers:
            h = layer(h)
        return h.squeeze(-1)


class _ExportPathCV(nn.Module):
    def __init__(self, degree: int, x_i, y_i, lam: float, weights, phi: float):
        super().__init__()
        self.degree, self.lam, self.phi = degree, lam, phi
        self.x_i, self.y_i = x_i, y_i
        self.weights = nn.Parameter(weights)

    def forward(self, positions):
        d = (positions[:, 0:1] - self.x_i.to(positions.dtype)) ** 2 + (positions[:, 1:2] - self.y_i.to(positions.dtype)) ** 2
        a = -self.lam * d
        idx = torch.arange(1, self.x_i.shape[0] + 1, dtype=positions.dtype, device=positions.device)
        l0 = torch.logsumexp(a, 1)
        s = torch.exp(torch.logsumexp(a + torch.log(idx), 1) - l0) / self.x_i.shape[0]
        table = _legendre_stack(torch.clamp(2.0 * s - 1.0, -1.0, 1.0), self.degree)
        return (self.weights.to(table.dtype).unsqueeze(1) * table).sum(0) - self.phi * l0 / self.lam
