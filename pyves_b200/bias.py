"""Bias potentials and their update rules -- mirror of ``ves/bias.py``.

Reference classes kept with the same names and signatures: ``Bias`` (``ves/bias.py:16-30``),
``NNBias`` (``:33-59``), ``HarmonicBias_SingleParticle_1D`` (+ ``_ForceModule``, ``:76-133``),
``StaticBias_SingleParticle`` (``:136-152``), ``VESBias_SingleParticle_1D`` (``:160-253``).

Differences in mechanism, not in behaviour:
* ``.force`` returns a ``BiasDescriptor`` (bias kind + parameters) that the simulation uploads to its
  CUDA handle, instead of an ``openmmtorch.TorchForce`` that re-reads ``model_loc`` from disk
  (``ves/bias.py:56-59``).  ``model_loc`` still receives the TorchScript export (``:52-53,249-253``).
* ``VESBias_SingleParticle_1D.update(traj)`` evaluates the reference's loss
  ``(1/beta) log sum_t exp(beta V(x_t)) + sum_i p_i V(x_i)`` (``:216-236``) and its parameter gradient
  with the CUDA kernels (``pyves_eval_bias`` + ``pyves_moments`` with soft-max row weights) instead of
  <= 5200 Python-level module calls and autograd; the optimiser is the same ``torch.optim`` object.

New: ``VESBias_Ensemble`` -- the textbook VES estimator (gradient ``-<f>_V + <f>_p``, Hessian
``beta Cov_V f``) fed by ensemble moments, with averaged SGD (1st / 2nd order) or a ``torch.optim``
optimiser, a refreshable (well-tempered) target and optional gradient running averages.
"""
import numpy as np
import torch
import torch.optim as optim

from . import _lib
from .basis import LegendreBasis1D, LegendreBasis2D, LegendreBasis2DPathCV, NNBasis1D, _KernelModule, _axis_id
import warnings

from .optim import AveragedSGD, EmptyEnsembleAverage, RunningAverage, gradient_hessian


class BiasDescriptor:
    """What ``Bias.force`` hands to the simulation: a bias kind and its parameters."""

    def __init__(self, kind, params, moment_inside=None):
        self.kind = kind
        self.params = params
        self.moment_inside = moment_inside     # None: leave the handle's moment domain as it is

    def apply(self, ens):
        p = self.params
        if self.moment_inside is not None:
            ens.set_moment_domain(self.moment_inside)
        if self.kind == "legendre1d":
            ens.set_bias_legendre1d(p["axis"], p["degree"], p["lo"], p["hi"], p["w"])
        elif self.kind == "legendre2d":
            ens.set_bias_legendre2d(p["deg_x"], p["deg_y"], p["minmax"], p["w"])
        elif self.kind == "mlp":
            ens.set_bias_mlp(p["axis"], p["lo"], p["hi"], p["sizes"], p["act"], p["theta"])
        elif self.kind == "pathcv":
            ens.set_bias_pathcv(p["degree"], p["x_i"], p["y_i"], p["lam"], p["w"], p["phi"])
        elif self.kind == "harmonic":
            ens.set_bias_harmonic(p["axis"], p["k"], p["s0"])
        elif self.kind == "none":
            ens.set_bias_none()
        else:
            raise ValueError("unknown bias kind %r" % self.kind)


################################################################################
# Base classes
################################################################################
class Bias:
    """Base class defining biases."""

    def __init__(self):
        pass

    def update(self, traj):
        # Biases are static by default.
        raise NotImplementedError("This bias cannot be updated.")

    @property
    def force(self):
        pass


class NNBias(Bias):
    """Base of the biases defined by a module ``self.model`` (set by the child class before
    ``super().__init__()``); exports it as TorchScript to ``model_loc``."""

    def __init__(self, model_loc="model.pt"):
        self.model_loc = model_loc
        self._export(self.model_loc)
        super().__init__()

    def _export(self, *paths):
        if self.model_loc is None:
            return
        module = self.model.export_torchscript()
        for path in paths:
            module.save(path)

    @property
    def force(self):
        return BiasDescriptor(*self.model.descriptor())


################################################################################
# Static biases
################################################################################
class HarmonicBias_SingleParticle_1D_ForceModule(_KernelModule):
    """Harmonic potential $k/2 (s - s0)^2$ along x or y (``ves/bias.py:97-133``); like upstream, the
    energies of all rows are summed into one scalar."""

    def __init__(self, k, s0, axis='x'):
        self.k = k
        self.s0 = s0
        self.axis = axis
        super().__init__()

    def _kernel_params(self):
        return ()

    def _upload(self, ens, params):
        ens.set_bias_harmonic(_axis_id(self.axis), float(self.k), float(self.s0))

    def forward(self, positions):
        _axis_id(self.axis)
        return torch.sum(self._run(positions))

    def descriptor(self):
        return ("harmonic", dict(axis=_axis_id(self.axis), k=float(self.k), s0=float(self.s0)))

    def export_torchscript(self):
        return torch.jit.script(_ExportHarmonic(float(self.k), float(self.s0), _axis_id(self.axis)))


class _ExportHarmonic(torch.nn.Module):
    def __init__(self, k: float, s0: float, axis: int):
        super().__init__()
        self.k, self.s0, self.axis = k, s0, axis

    def forward(self, positions):
        return 0.5 * self.k * torch.sum((positions[:, self.axis] - self.s0) ** 2)


class HarmonicBias_SingleParticle_1D(NNBias):
    """Harmonic bias $k/2 (s - s0)^2$ along the x- or y-coordinate (``ves/bias.py:76-94``)."""

    def __init__(self, k, s0, axis='x', model_loc="model.pt"):
        if axis not in ('x', 'y'):
            raise ValueError("Invalid axis")
        self.model = HarmonicBias_SingleParticle_1D_ForceModule(k, s0, axis=axis)
        super().__init__(model_loc)


class StaticBias_SingleParticle(NNBias):
    """A static basis-set expanded potential: ``V_module`` is any module of ``pyves_b200.basis``
    (``ves/bias.py:136-152``)."""

    def __init__(self, V_module, model_loc):
        self.model = V_module
        super().__init__(model_loc)


################################################################################
# Dynamic biases
################################################################################
def _make_torch_optimizer(optimizer_type, params, optimizer_params):
    if optimizer_type == 'SGD':
        return optim.SGD(params, **optimizer_params)
    if optimizer_type == 'RMSprop':
        return optim.RMSprop(params, **optimizer_params)
    if optimizer_type == 'Adam':
        return optim.Adam(params, **optimizer_params)
    raise ValueError("Requested optimizer not yet supported.")


def _set_grads(model, flat):
    flat = torch.as_tensor(flat, dtype=torch.float64)
    o = 0
    for p in model.parameters():
        p.grad = flat[o:o + p.numel()].reshape(p.shape).to(device=p.device, dtype=p.dtype).clone()
        o += p.numel()


def _device_eval(model, pos_np):
    """Energies of rows ``pos_np`` [N, C] (float64) through the module's CUDA evaluator, no autograd."""
    t = torch.as_tensor(np.ascontiguousarray(pos_np, dtype=np.float64), device=torch.device("cuda", torch.cuda.current_device()))
    ens, _ = model._evaluator(t)
    model._upload(ens, model._kernel_params())
    return ens, t


class VESBias_SingleParticle_1D(NNBias):
    """Dynamic basis-set expanded potential along x or y of a single particle; ``update(traj)`` makes
    one optimiser step on the reference's loss over the last ``update_steps`` frames
    (``ves/bias.py:160-253``).

    ``V_module`` is a ``LegendreBasis1D`` or ``NNBasis1D``; ``target`` a ``pyves_b200.target.Target_x``.
    """

    def __init__(self, V_module, target, beta, optimizer_type, optimizer_params, model_loc, update_steps=5000):
        self.model = V_module
        self.target = target
        self.beta = beta
        self.optimizer = _make_torch_optimizer(optimizer_type, self.model.parameters(), optimizer_params)
        self.update_steps = update_steps
        super().__init__(model_loc)

    def loss_and_grad(self, traj):
        """(loss, flat parameter gradient) of the reference's loss, computed on the GPU in FP64."""
        traj = np.asarray(traj, dtype=np.float64)
        T = traj.shape[0]
        lo = max(T - self.update_steps, 0)
        print(lo, T)                                                # ves/bias.py:218
        ens, win = _device_eval(self.model, traj[lo:T].reshape(T - lo, -1))
        E, _ = ens.eval_bias(win, want_force=False)
        bE = self.beta * E
        m = torch.max(bE)
        ex = torch.exp(bE - m)
        z = torch.sum(ex)
        loss_V = (m + torch.log(z)) / self.beta
        _, g_traj, _ = ens.moments(pos=win, row_weights=(ex / z).contiguous())
        # target nodes are fed through the module as raw coordinates (ves/bias.py:229-233)
        a = _axis_id(self.model.axis)
        nodes = np.zeros((len(self.target.x), a + 1))
        nodes[:, a] = self.target.x
        tpos = torch.as_tensor(nodes, device="cuda")
        p = torch.as_tensor(np.ascontiguousarray(self.target.p, dtype=np.float64), device="cuda")
        Et, _ = ens.eval_bias(tpos, want_force=False)
        _, g_tgt, _ = ens.moments(pos=tpos, row_weights=p)
        loss = float(loss_V + torch.dot(p, Et))
        return loss, (g_traj + g_tgt).cpu()

    def update(self, traj):
        self.optimizer.zero_grad()
        loss, grad = self.loss_and_grad(traj)
        print("Loss = {:.6f}".format(loss))
        _set_grads(self.model, grad)
        self.optimizer.step()
        # overwrite the current model and archive it (ves/bias.py:249-253)
        self._export(self.model_loc, self.model_loc + ".iter{}".format(np.asarray(traj).shape[0]))


class VESBias_Ensemble(NNBias):
    """Textbook VES update driven by ensemble moments (new; see module docstring).

    Args:
        V_module: ``LegendreBasis1D`` / ``LegendreBasis2D`` / ``LegendreBasis2DPathCV`` / ``NNBasis1D``.
        target: a target of ``pyves_b200.target`` whose nodes live in the scaled coordinate [-1, 1]
            of ``V_module`` (``Target_x`` family for 1D bases, ``Target_2D`` family for ``LegendreBasis2D``).
        beta: 1 / kT.
        optimizer_type: ``'averaged_sgd'`` (Bach-Moulines, ``optimizer_params`` = ``{'mu': .., 'second_order': bool,
            'averaging': 'ewma' | 'running', 'decay': 0.9, 'precondition': p}``; p > 0 scales the step of Legendre
            coefficient k by (2k+1)^p, the inverse variance of P_k under the flat target)
            or ``'SGD'`` / ``'RMSprop'`` / ``'Adam'`` (``torch.optim`` on the textbook gradient).
        model_loc: TorchScript export path or None.
        target_update_every: refresh a well-tempered target every this many updates.
        gradient_average: None, ``'running'`` or a decay in (0, 1) for an exponentially weighted mean.
        bias_average: None or a decay d in (0, 1): keep ``V_avg <- d V_avg + (1 - d) (V - mean V)`` of the bias at the
            target nodes over the updates and read the free energy from it (``fes()``).  Averaged SGD applies and reads
            the AVERAGED coefficients; a ``torch.optim`` optimiser on an NN has no such average, its bias jitters around
            the optimum with the step size, and the update-averaged bias is the estimator with the same role
            (measured on configs[3], ``profiles/r02_nn_ves_convergence.md``: 0.03-0.05 kT against 0.13-0.15 kT for the
            instantaneous bias).
        sample_domain: ``'box'`` (default) -- <f>_V is the average over the walkers INSIDE the basis interval
            (the support of the target), so walkers that wander outside, where the scaled coordinate clamps
            (``ves/basis.py:143``), do not pile up on the edge of the histogram the bias is matched to;
            ``'all'`` -- every walker counts with its clamped coordinate.  Measured on configs[1]
            (double well, 16 x 16, gamma 10, box y in [-2, 2] where F is only 8 kT): converged 2D FES RMSE
            0.03 kT with ``'box'`` against a plateau of 0.14 kT with ``'all'``.
    """

    def __init__(self, V_module, target, beta, optimizer_type='averaged_sgd', optimizer_params=None, model_loc=None,
                 target_update_every=10, gradient_average=None, sample_domain='box', bias_average=None):
        if sample_domain not in ('box', 'all'):
            raise ValueError("sample_domain must be 'box' or 'all'")
        self.sample_domain = sample_domain
        self.model = V_module
        self.target = target
        self.beta = beta
        optimizer_params = dict(optimizer_params or {})
        self.optimizer_type = optimizer_type
        if optimizer_type == 'averaged_sgd':
            if isinstance(V_module, NNBasis1D):
                raise ValueError("averaged_sgd needs a linear basis expansion; use Adam / SGD / RMSprop for NNBasis1D")
            w = np.concatenate([p.detach().cpu().numpy().ravel() for p in V_module.parameters()])
            scale = None
            pw = float(optimizer_params.get('precondition', 0.0))      # exponent of the diagonal preconditioner
            if pw:
                if isinstance(V_module, LegendreBasis2D):
                    scale = np.outer(2 * np.arange(V_module.degree_x + 1) + 1, 2 * np.arange(V_module.degree_y + 1) + 1).ravel()
                else:
                    scale = 2 * np.arange(V_module.degree + 1) + 1
                scale = scale.astype(np.float64) ** pw
            self.optimizer = AveragedSGD(w, mu=optimizer_params.get('mu', 1.0),
                                         second_order=optimizer_params.get('second_order', False),
                                         averaging=optimizer_params.get('averaging', 'ewma'),
                                         decay=optimizer_params.get('decay', 0.9), scale=scale)
        else:
            self.optimizer = _make_torch_optimizer(optimizer_type, self.model.parameters(), optimizer_params)
        self.target_update_every = int(target_update_every)
        if gradient_average is None:
            self.grad_avg = None
        else:
            self.grad_avg = RunningAverage(None if gradient_average == 'running' else float(gradient_average))
        if bias_average is not None and not 0.0 < float(bias_average) < 1.0:
            raise ValueError("bias_average must be a decay in (0, 1)")
        self.bias_average = None if bias_average is None else float(bias_average)
        self._V_avg = None               # exponentially weighted mean over the updates of the (mean-free) bias at the target nodes
        self.n_updates = 0
        self._f_target = None
        self.last_gradient = None
        super().__init__(model_loc)

    # -------------------------------------------------------------------------------------------
    @property
    def force(self):
        kind, params = self.model.descriptor()
        return BiasDescriptor(kind, params, moment_inside=(self.sample_domain == 'box'))

    @property
    def needs_covariance(self):
        return self.optimizer_type == 'averaged_sgd' and self.optimizer.second_order

    def _target_positions(self):
        """Target nodes mapped from the scaled coordinate to real coordinates, [M, C]."""
        m = self.model
        s = np.asarray(self.target.x, dtype=np.float64)
        if isinstance(m, LegendreBasis2D):
            if getattr(self.target, "ndim", 1) != 2:
                raise ValueError("LegendreBasis2D needs a 2D target")
            cx, hx = (m.min_x + m.max_x) / 2, (m.max_x - m.min_x) / 2
            cy, hy = (m.min_y + m.max_y) / 2, (m.max_y - m.min_y) / 2
            return np.stack([cx + hx * s[:, 0], cy + hy * s[:, 1]], axis=1)
        if isinstance(m, LegendreBasis2DPathCV):
            raise NotImplementedError("path-CV targets live in the scaled path coordinate; see _pathcv_nodes")
        a = _axis_id(m.axis)
        pos = np.zeros((len(s), a + 1))
        if isinstance(m, NNBasis1D):
            pos[:, a] = m.min + (m.max - m.min) * (s + 1) / 2      # NN scaling maps [min, max] to [0, 1]
        else:
            pos[:, a] = (m.min + m.max) / 2 + (m.max - m.min) / 2 * s
        return pos

    # ---- path CV: the target is a 1D target in the scaled path coordinate s' = 2 s - 1; f_k(s') = P_k(s') does not
    # depend on where on the plane a node sits, so node averages and the bias at the nodes are plain Legendre sums
    def _pathcv_nodes(self):
        if getattr(self.target, "ndim", 1) != 1:
            raise ValueError("LegendreBasis2DPathCV needs a 1D target (nodes in the scaled path coordinate)")
        return np.polynomial.legendre.legvander(np.asarray(self.target.x, dtype=np.float64), self.model.degree)   # [M, K]

    def _pathcv_bias_at_nodes(self):
        w = self.model.weights.detach().cpu().numpy().astype(np.float64)
        return self._pathcv_nodes() @ w

    def target_average(self, refresh=False):
        """<dV/dparam>_p by quadrature over the target nodes (CUDA moments kernel with node weights)."""
        if isinstance(self.model, LegendreBasis2DPathCV):
            if self._f_target is None or refresh:
                self._f_target = self.target.quadrature_weights() @ self._pathcv_nodes()
            return self._f_target
        if self._f_target is None or refresh or isinstance(self.model, NNBasis1D):
            ens, tpos = _device_eval(self.model, self._target_positions())
            q = torch.as_tensor(self.target.quadrature_weights(), device="cuda")
            sw, sf, _ = ens.moments(pos=tpos, row_weights=q.contiguous())
            self._f_target = (sf / sw).cpu().numpy()
        return self._f_target

    def refresh_target(self):
        if isinstance(self.model, LegendreBasis2DPathCV):
            if self.target.update(self._pathcv_bias_at_nodes(), self.beta):
                self._f_target = None
            return
        ens, tpos = _device_eval(self.model, self._target_positions())
        V, _ = ens.eval_bias(tpos, want_force=False)
        if self.target.update(V.cpu().numpy(), self.beta):
            self._f_target = None

    def update_from_moments(self, sum_w, sum_wf, sum_wff=None):
        """One optimiser step from ensemble sums (already reduced over all GPUs)."""
        tonp = lambda t: t.detach().cpu().numpy() if torch.is_tensor(t) else np.asarray(t)
        sum_w = float(tonp(sum_w).ravel()[0])
        f_p = self.target_average()
        try:
            grad, hess = gradient_hessian(sum_w, tonp(sum_wf), None if sum_wff is None else tonp(sum_wff), f_p, self.beta)
        except EmptyEnsembleAverage as exc:
            # same rule as the device kernel (averaged_sgd_kernel): an update without samples changes nothing
            warnings.warn("VES update skipped: %s" % exc)
            if self.optimizer_type == 'averaged_sgd':
                self.optimizer.n += 1
            self.n_updates += 1
            return
        if self.grad_avg is not None:
            grad = self.grad_avg.update(grad)
        self.last_gradient = grad
        if self.optimizer_type == 'averaged_sgd':
            abar = self.optimizer.step(grad, hess)
            with torch.no_grad():
                o = 0
                for p in self.model.parameters():
                    p.copy_(torch.from_numpy(abar[o:o + p.numel()].reshape(tuple(p.shape))))
                    o += p.numel()
        else:
            self.optimizer.zero_grad()
            _set_grads(self.model, grad)
            self.optimizer.step()
        self.n_updates += 1
        if self.bias_average is not None:
            V = self._bias_at_nodes()
            V = V - V.mean()
            self._V_avg = V if self._V_avg is None else self.bias_average * self._V_avg + (1.0 - self.bias_average) * V
        if self.target_update_every > 0 and self.n_updates % self.target_update_every == 0:
            self.refresh_target()
        if self.model_loc is not None:
            self._export(self.model_loc, self.model_loc + ".iter{}".format(self.n_updates))

    def update(self, traj):
        """Reference-style entry point: the rows of ``traj`` [T, C] are the samples of <.>_V."""
        ens, rows = _device_eval(self.model, np.asarray(traj, dtype=np.float64).reshape(len(traj), -1))
        ens.set_moment_domain(self.sample_domain == 'box')
        try:
            sw, sf, sff = ens.moments(pos=rows, cov=self.needs_covariance)
        finally:
            ens.set_moment_domain(False)       # the evaluator is shared with the module's other users
        self.update_from_moments(sw, sf, sff)

    # ---- checkpoint of the whole update state (the reference archives only the TorchScript module,
    # ves/bias.py:249-253, and pickles positions / velocities, ves/langevin_dynamics.py:197-204)
    def state_dict(self):
        """Everything a resumed run needs to continue the optimisation exactly: parameters, optimiser state (averaged-SGD
        iterates and averages, or the ``torch.optim`` state), gradient running average, target density, update count."""
        d = dict(optimizer_type=self.optimizer_type, n_updates=int(self.n_updates),
                 model=[p.detach().cpu().numpy().copy() for p in self.model.parameters()],
                 target_p=np.array(self.target.p, dtype=np.float64, copy=True),
                 f_target=None if self._f_target is None else np.array(self._f_target, dtype=np.float64, copy=True),
                 sgd_steps=int(getattr(self, "_sgd_steps", 0)),
                 V_avg=None if self._V_avg is None else np.array(self._V_avg, dtype=np.float64, copy=True))
        if self.optimizer_type == 'averaged_sgd':
            d["optimizer"] = self.optimizer.state_dict()
        else:
            d["optimizer"] = self.optimizer.state_dict()
        if self.grad_avg is not None:
            d["grad_avg"] = dict(decay=self.grad_avg.decay, n=int(self.grad_avg.n),
                                 value=None if self.grad_avg.value is None else np.array(self.grad_avg.value, copy=True))
        return d

    def load_state_dict(self, d):
        if d.get("optimizer_type") != self.optimizer_type:
            raise ValueError("checkpoint was written by optimizer %r, this bias uses %r" % (d.get("optimizer_type"), self.optimizer_type))
        params = list(self.model.parameters())
        if len(params) != len(d["model"]) or any(tuple(p.shape) != tuple(a.shape) for p, a in zip(params, d["model"])):
            raise ValueError("checkpointed parameters do not match the model")
        with torch.no_grad():
            for p, a in zip(params, d["model"]):
                p.copy_(torch.from_numpy(np.asarray(a)))
        self.optimizer.load_state_dict(d["optimizer"])
        self.n_updates = int(d["n_updates"])
        self._sgd_steps = int(d.get("sgd_steps", 0))
        if len(d["target_p"]) != len(self.target.p):
            raise ValueError("checkpointed target density does not match the target mesh")
        self.target._p = np.array(d["target_p"], dtype=np.float64, copy=True)
        self._f_target = None if d.get("f_target") is None else np.array(d["f_target"], dtype=np.float64, copy=True)
        self._V_avg = None if d.get("V_avg") is None else np.array(d["V_avg"], dtype=np.float64, copy=True)
        if self.grad_avg is not None and "grad_avg" in d:
            self.grad_avg.n = int(d["grad_avg"]["n"])
            v = d["grad_avg"]["value"]
            self.grad_avg.value = None if v is None else np.array(v, dtype=np.float64, copy=True)

    def _bias_at_nodes(self):
        """Bias energy at the target nodes (numpy, kJ/mol)."""
        if isinstance(self.model, LegendreBasis2DPathCV):
            return self._pathcv_bias_at_nodes()
        ens, tpos = _device_eval(self.model, self._target_positions())
        V, _ = ens.eval_bias(tpos, want_force=False)
        return V.cpu().numpy()

    def fes(self, nodes_scaled=None, averaged=None):
        """F(s) = -V(s) - (1/beta) ln p(s) at the target nodes, shifted to min 0 (kJ/mol).  ``averaged``: read V from the
        update-averaged bias (default: whenever ``bias_average`` is set and at least one update was made)."""
        if averaged is None:
            averaged = self._V_avg is not None
        if averaged and self._V_avg is None:
            raise ValueError("no averaged bias: construct the bias with bias_average=decay and run updates")
        V = self._V_avg if averaged else self._bias_at_nodes()
        F = -V - np.log(np.maximum(self.target.p, 1e-300)) / self.beta
        return F - F.min()


def _torch_optimizer_plan(opt):
    """(kind, hyper, why_not) of a ``torch.optim`` optimiser for ``pyves_optimizer_step``: the plain SGD / RMSprop / Adam
    rules the reference constructs (``ves/bias.py:193-200``); ``why_not`` names the option the kernel does not implement."""
    if len(opt.param_groups) != 1:
        return None, None, "several parameter groups"
    g = opt.param_groups[0]
    if g.get("weight_decay", 0) or g.get("maximize", False):
        return None, None, "weight_decay / maximize"
    if isinstance(opt, optim.Adam):
        if g.get("amsgrad", False) or type(opt) is not optim.Adam:
            return None, None, "amsgrad / Adam subclass"
        return "Adam", (g["lr"], g["betas"][0], g["betas"][1], g["eps"]), None
    if isinstance(opt, optim.RMSprop):
        if g.get("momentum", 0) or g.get("centered", False):
            return None, None, "RMSprop momentum / centered"
        return "RMSprop", (g["lr"], g["alpha"], 0.0, g["eps"]), None
    if isinstance(opt, optim.SGD):
        if g.get("nesterov", False) or g.get("dampening", 0):
            return None, None, "SGD nesterov / dampening"
        return "SGD", (g["lr"], g.get("momentum", 0.0), 0.0, 0.0), None
    return None, None, "optimizer type"


class DeviceUpdater:
    """Device-resident update loop of a ``VESBias_Ensemble``: the optimiser state, the target density, the target
    averages and the bias parameters live in CUDA tensors, one update is a handful of launches enqueued on the
    current stream, and the new parameters reach the step kernels through ``pyves_set_bias_*_device`` -- no
    device-to-host read, no host optimiser, no re-upload between two launches, so the GPU never waits for Python.

    * ``LegendreBasis1D`` / ``LegendreBasis2D`` with ``averaged_sgd`` (first or second order): gradient, Hessian term,
      preconditioner and averaging are ONE kernel (``pyves_averaged_sgd_step``);
    * ``NNBasis1D`` (two or more hidden layers) with ``SGD`` / ``RMSprop`` / ``Adam`` and the optional running /
      exponentially weighted gradient average: ONE kernel (``pyves_optimizer_step``); the parameter-gradient target
      average <dV/dtheta>_p is re-evaluated with the new parameters after every step (``pyves_moments`` on the
      target nodes), the fold + repack of the layers for the step kernels is ``pyves_set_bias_mlp_device``.

    Numerically it is the same update as ``VESBias_Ensemble.update_from_moments`` (FP64), which remains the reference
    implementation; ``sync_to_host()`` copies parameters, optimiser state and target back into the bias object (and
    its ``torch.optim`` optimiser), from which a new ``DeviceUpdater`` resumes.

    Replaces, for the ensemble, the per-update path ``bias.update`` -> TorchScript save -> ``TorchForce`` reload
    -> ``context.reinitialize`` of ``ves/langevin_dynamics.py:147-159`` + ``ves/bias.py:206-253``.
    """

    @staticmethod
    def why_not(bias):
        """None when ``bias`` can be updated on the device, else the reason."""
        if not isinstance(bias, VESBias_Ensemble):
            return "not a VESBias_Ensemble"
        m = bias.model
        if isinstance(m, (LegendreBasis1D, LegendreBasis2D)):
            if bias.optimizer_type != 'averaged_sgd':
                return "Legendre bases are updated on the device with averaged_sgd only"
            if bias.grad_avg is not None:
                return "averaged_sgd on the device has no gradient averaging"
            return None
        if isinstance(m, NNBasis1D):
            if bias.optimizer_type == 'averaged_sgd':
                return "averaged_sgd needs a linear basis"
            if len(m.hidden_layer_sizes) < 2:
                return "a single hidden layer is the linear case (host path)"
            return _torch_optimizer_plan(bias.optimizer)[2]
        return "basis %s is updated on the host" % type(m).__name__

    @staticmethod
    def supports(bias):
        return DeviceUpdater.why_not(bias) is None

    def __init__(self, bias, ens):
        why = DeviceUpdater.why_not(bias)
        if why is not None:
            raise ValueError("DeviceUpdater: " + why)
        m = bias.model
        self.bias, self.ens = bias, ens
        self.dev = torch.device("cuda", ens.device)
        self.nn = isinstance(m, NNBasis1D)
        self.two_d = isinstance(m, LegendreBasis2D)
        f64 = dict(dtype=torch.float64, device=self.dev)
        self.n_updates = bias.n_updates
        # target: nodes in real coordinates, density, node averages of the basis functions (FP64 evaluator handle)
        self.ev = _lib.Ensemble(1, dims=2, precision="fp64", device=ens.device)
        self.tpos = torch.tensor(bias._target_positions(), **f64).contiguous()
        self.p = torch.tensor(np.asarray(bias.target.p, dtype=np.float64), **f64)
        self.refreshable = hasattr(bias.target, "bias_factor")
        if getattr(bias.target, "ndim", 1) == 2:
            self.cell = 4.0 / float(len(bias.target.p))
        else:
            x = np.asarray(bias.target.x, dtype=np.float64)
            self.dx = float(x[1] - x[0])
            tw = np.ones(len(x))
            tw[0] = tw[-1] = 0.5                               # trapezoid rule of Target_x.quadrature_weights
            self.trap = torch.tensor(tw, **f64)
        self.handles = [ens]                 # every handle that receives the parameters (add_handle)
        self._flat = None
        self.V_avg = None if bias._V_avg is None else torch.tensor(bias._V_avg, **f64)
        if self.nn:
            self._init_nn(f64)
        else:
            o = bias.optimizer
            self.alpha = torch.tensor(o.alpha, **f64)
            self.abar = torch.tensor(o.abar, **f64)
            self.scale = None if o.scale is None else torch.tensor(o.scale, **f64)
            self.mu, self.decay, self.averaging, self.n = o.mu, o.decay, o.averaging, o.n
            self.second_order = bool(o.second_order)
            self._spare = (torch.empty_like(self.alpha), torch.empty_like(self.abar))
        self._upload(self.ev)
        self.f_target = self._target_average()
        ens.set_moment_domain(bias.sample_domain == 'box')
        self._upload(ens)
        if self.refreshable:
            self._refreshed_density()       # result discarded: loads the kernels of the refresh path (CUDA loads lazily)

    # ---- NN state: flat parameter vector + torch.optim state + gradient average, all FP64 on the device
    def _init_nn(self, f64):
        b, m = self.bias, self.bias.model
        self.second_order = False
        params = list(m.parameters())
        self.theta = torch.cat([p.detach().reshape(-1).to(torch.float64) for p in params]).to(self.dev).contiguous()
        P = self.theta.numel()
        self.kind, self.hyper, _ = _torch_optimizer_plan(b.optimizer)
        self.s1 = torch.zeros(P, **f64)
        self.s2 = torch.zeros(P, **f64)
        self.t = 0                                           # optimiser steps done
        st = b.optimizer.state
        if len(st) > 0:                                      # resume from the host optimiser's state
            key1 = {"Adam": "exp_avg", "RMSprop": "square_avg", "SGD": "momentum_buffer"}[self.kind]
            cat = lambda key: torch.cat([(st[p][key] if st[p].get(key) is not None else torch.zeros_like(p)).detach().reshape(-1).to(torch.float64)
                                         for p in params]).to(self.dev).contiguous()
            self.s1 = cat(key1)
            if self.kind == "Adam":
                self.s2 = cat("exp_avg_sq")
            step = st[params[0]].get("step", None)
            self.t = int(step) if step is not None else int(getattr(b, "_sgd_steps", 0))
        ga = b.grad_avg
        self.gavg = None if ga is None else ("running" if ga.decay is None else "ewma")
        self.gdecay = 0.0 if (ga is None or ga.decay is None) else float(ga.decay)
        self.gn = 0 if ga is None else int(ga.n)
        self.gstate = torch.zeros(P, **f64)
        if ga is not None and ga.value is not None:
            self.gstate.copy_(torch.as_tensor(ga.value, **f64))

    @property
    def K(self):
        """Number of basis functions / parameters (length of ``sum_wf``)."""
        return (self.theta if self.nn else self.alpha).numel()

    @property
    def lr(self):
        return self.hyper[0]

    @lr.setter
    def lr(self, value):
        """Learning-rate schedules are host arithmetic: the rate is a by-value argument of every step."""
        self.hyper = (float(value),) + tuple(self.hyper[1:])

    def add_handle(self, ens):
        """Another ensemble handle on the same device (a shard processed on its own stream) that shares the bias.
        ``update`` hands the new coefficients to every handle on the stream it is called on; a shard stepped on
        another stream must wait for that stream first (``shard_stream.wait_stream(update_stream)``, as bench.py's
        chunked end-to-end loop does)."""
        ens.set_moment_domain(self.bias.sample_domain == 'box')
        self._upload(ens)
        self.handles.append(ens)

    def _upload(self, handle):
        m = self.bias.model
        if self.nn:
            handle.set_bias_mlp_device(_axis_id(m.axis), float(m.min), float(m.max), m.hidden_layer_sizes, m._act, self.theta)
            return
        if handle is not self.ens and handle is not self.ev and handle.precision == self.ens.precision:
            handle.set_bias_like(self.ens)          # shards of one ensemble share the converted weights (and their tag)
            return
        if self.two_d:
            handle.set_bias_legendre2d_device(m.degree_x, m.degree_y, (m.min_x, m.max_x, m.min_y, m.max_y), self.abar)
        else:
            handle.set_bias_legendre1d_device(_axis_id(m.axis), m.degree, float(m.min), float(m.max), self.abar)

    def _target_average(self):
        """<dV/dparam>_p by quadrature over the target nodes (the evaluator handle must hold the current parameters)."""
        q = self.p if getattr(self.bias.target, "ndim", 1) == 2 else self.p * self.trap   # Target_2D / Target_x .quadrature_weights
        q = (q / q.sum()).contiguous()
        sw, sf, _ = self.ev.moments(pos=self.tpos, row_weights=q)
        return sf / sw

    def update(self, sum_w, sum_wf, sum_wff=None):
        """One update from ensemble sums on the device (already reduced over all GPUs).  Enqueue only."""
        with _lib._Range("pyves.update"):
            if self.nn:
                self._update_nn(sum_w, sum_wf)
            else:
                self._update(sum_w, sum_wf, sum_wff)

    def _update_nn(self, sum_w, sum_wf):
        self.t += 1
        if self.gavg is not None:
            self.gn += 1
        _lib.optimizer_step(self.ens.device, self.kind, sum_w, sum_wf, self.f_target, self.gavg, self.gdecay, max(self.gn, 1),
                            self.gstate if self.gavg is not None else None, self.theta, self.s1, self.s2 if self.kind == "Adam" else None,
                            self.hyper, self.t)
        self.n_updates += 1
        for hdl in self.handles:
            self._upload(hdl)
        self._upload(self.ev)
        self._average_bias(uploaded=True)
        b = self.bias
        if self.refreshable and b.target_update_every > 0 and self.n_updates % b.target_update_every == 0:
            self.p = self._refreshed_density()
        self.f_target = self._target_average()              # <dV/dtheta>_p depends on theta: for the next update

    def _update(self, sum_w, sum_wf, sum_wff):
        if self.second_order and sum_wff is None:
            raise ValueError("the second-order step needs the second moments (ens.moments(cov=True))")
        self.n += 1
        a_out, b_out = self._spare                                    # ping-pong: the kernel must not update in place
        _lib.averaged_sgd_step(self.ens.device, sum_w, sum_wf, sum_wff if self.second_order else None, self.f_target, self.alpha,
                               self.abar, self.scale, self.mu, self.bias.beta, self.averaging, self.decay, self.n, a_out, b_out)
        self._spare = (self.alpha, self.abar)
        self.alpha, self.abar = a_out, b_out
        self.n_updates += 1
        for hdl in self.handles:
            self._upload(hdl)
        self._average_bias(uploaded=False)
        b = self.bias
        if self.refreshable and b.target_update_every > 0 and self.n_updates % b.target_update_every == 0:
            self.p = self._refreshed_density()
            self.f_target = self._target_average()

    def _average_bias(self, uploaded):
        """VESBias_Ensemble's ``bias_average`` on the device: EWMA over the updates of the mean-free bias at the target nodes."""
        d = self.bias.bias_average
        if d is None:
            return
        if not uploaded:
            self._upload(self.ev)
        V, _ = self.ev.eval_bias(self.tpos, want_force=False)
        V = V - V.mean()
        self.V_avg = V if self.V_avg is None else d * self.V_avg + (1.0 - d) * V

    def moment_buffers(self, cov=None):
        """Views (sum_w, sum_wf, sum_wff or None) of ONE persistent flat FP64 buffer for ``ens.moments(out=...)``:
        no allocation per update, and ``reduce_flat`` all-reduces the flat buffer in place (one NCCL call)."""
        cov = self.second_order if cov is None else cov
        K = self.K
        n = 1 + K + (K * K if cov else 0)
        if getattr(self, "_flat", None) is None or self._flat.numel() != n:
            self._flat = torch.zeros(n, dtype=torch.float64, device=self.dev)
        f = self._flat
        return f[:1], f[1:1 + K], (f[1 + K:].view(K, K) if cov else None)

    def reduce_flat(self, group=None, comm=None):
        """Sum the flat moment buffer over the ranks in place: through the library's own collective
        (``comm``: a ``pyves_b200._lib.Communicator``, ``pyves_allreduce_moments``) or through ``torch.distributed``."""
        if comm is not None:
            if comm.nranks > 1:
                comm.allreduce_(self._flat)
            return
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self._flat, op=dist.ReduceOp.SUM, group=group)

    def _refreshed_density(self):
        """Target_WellTempered_{x,2D}.update on the device: p' ~ exp(-(beta / gamma) F), F = -V - kT ln p."""
        b = self.bias
        self._upload(self.ev)
        V, _ = self.ev.eval_bias(self.tpos, want_force=False)
        F = -V - torch.log(torch.clamp(self.p, min=1e-300)) / b.beta
        q = torch.exp(-(b.beta / b.target.bias_factor) * (F - F.min()))
        if getattr(b.target, "ndim", 1) == 2:
            return q / (q.sum() * self.cell)
        return q / (torch.sum(q * self.trap) * self.dx)

    def bias_at_nodes(self):
        """Bias energy at the target nodes with the current device parameters (CUDA float64 tensor [M])."""
        self._upload(self.ev)
        V, _ = self.ev.eval_bias(self.tpos, want_force=False)
        return V

    def sync_to_host(self):
        """Copy the device state back into the bias object (parameters, optimiser, gradient average, target)."""
        b, o = self.bias, self.bias.optimizer
        b.n_updates = self.n_updates
        if self.nn:
            theta = self.theta.cpu()
            s1, s2, g = self.s1.cpu(), self.s2.cpu(), self.gstate.cpu()
            with torch.no_grad():
                off = 0
                for prm in b.model.parameters():
                    n = prm.numel()
                    prm.copy_(theta[off:off + n].reshape(prm.shape))
                    if self.t > 0:
                        st = o.state[prm]
                        if self.kind == "Adam":
                            st["step"] = torch.tensor(float(self.t))
                            st["exp_avg"] = s1[off:off + n].reshape(prm.shape).clone()
                            st["exp_avg_sq"] = s2[off:off + n].reshape(prm.shape).clone()
                        elif self.kind == "RMSprop":
                            st["step"] = torch.tensor(float(self.t))
                            st["square_avg"] = s1[off:off + n].reshape(prm.shape).clone()
                        elif self.hyper[1] != 0.0:
                            st["momentum_buffer"] = s1[off:off + n].reshape(prm.shape).clone()
                    off += n
            o.param_groups[0]["lr"] = self.hyper[0]
            b._sgd_steps = self.t
            if b.grad_avg is not None:
                b.grad_avg.n = self.gn
                b.grad_avg.value = g.numpy().copy() if self.gn > 0 else None
        else:
            o.alpha = self.alpha.cpu().numpy().copy()
            o.abar = self.abar.cpu().numpy().copy()
            o.n = self.n
            with torch.no_grad():
                off = 0
                for prm in b.model.parameters():
                    prm.copy_(torch.from_numpy(o.abar[off:off + prm.numel()].reshape(tuple(prm.shape))))
                    off += prm.numel()
        if self.refreshable:
            b.target._p = self.p.cpu().numpy().copy()
        b._f_target = self.f_target.cpu().numpy().copy()
        b._V_avg = None if self.V_avg is None else self.V_avg.cpu().numpy().copy()

    def close(self):
        self.ev.close()
