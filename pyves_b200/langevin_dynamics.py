"""Biased Langevin dynamics driver -- mirror of ``ves/langevin_dynamics.py`` for 1..N walkers.

``SingleParticleSimulation`` keeps the reference's constructor, ``init_ves`` and ``__call__``
signatures, defaults, error messages, public attributes (``traj``, ``PE``, ``KE``) and output files
(``traj.dat`` / ``energies.dat`` formats of ``ves/langevin_dynamics.py:206-235``, a pickled state at
every checkpoint).  Instead of an OpenMM ``Context`` with one particle stepped one ``step(1)`` at a
time from Python (``:123,188``), the state of ``n_walkers`` independent walkers lives on the GPU and
the fused step kernel of the CUDA library advances whole stretches between two "events"
(trajectory frame, energy sample, checkpoint, bias update) in one launch.

New keywords: ``n_walkers`` (ensemble size; sharded over the ranks of an initialised
``torch.distributed`` process group), ``device``, ``precision`` ('fp32' | 'fp64'), ``scheme``
('openmm_langevin' | 'middle'), ``dims`` (2 drops the decoupled z coordinate, 3 keeps it),
``communicator`` (a ``pyves_b200._lib.Communicator``: the device-resident update loop sums the moments through the
library's own NCCL entry point ``pyves_allreduce_moments`` instead of ``torch.distributed``),
``traj_walkers`` (how many walkers are recorded), ``sample_every`` (moment sampling stride of the
ensemble VES update), ``device_update`` (keep optimiser state and coefficients of a ``VESBias_Ensemble`` on the
GPU between updates, ``bias.DeviceUpdater``; the bias object is synchronised when the run returns and at every
checkpoint; the per-update ``model_loc.iter{n}`` archive is written only with ``archive_updates=True``, which costs a
synchronisation per update).  Checkpoints of a dynamic ``VESBias_Ensemble`` run carry the whole update state
(``State``), and a run resumed from one continues the update cadence on the global loop index.  With ``n_walkers=1`` the behaviour and file contents follow the reference:
frame ``t`` is the position before step ``t`` (``:180-188``), ``bias.update(self.traj)`` runs at
``i == startafter`` and then whenever ``i % learnevery == 0`` (``:134-159``).
"""
import os
import pickle

import numpy as np
import torch

from . import _lib
from .bias import Bias, VESBias_Ensemble
from .bias import DeviceUpdater
from .parallel import dist_info, reduce_moments, shard

KB = 8.31446261815324e-3      # kJ/mol/K as OpenMM defines it (the integrator's kT)


class State:
    """Picklable snapshot (replaces ``openmm.State``): positions / velocities [n_walkers, dims], the Philox step
    counter and seed, and -- for a dynamic ``VESBias_Ensemble`` -- the whole update state (``bias.state_dict()``:
    coefficients / NN parameters, optimiser iterates and averages, gradient running average, target density, update
    count), the loop index of the run and of its last update, and the moment sums sampled since that update.  A run
    resumed from it (``init_state=``, then ``init_ves`` with a bias of the same construction) continues the
    uninterrupted run exactly (tested bit for bit in FP64)."""

    def __init__(self, positions, velocities, step, seed, walker_offset=0, loop_step=0, last_update_step=None, bias_state=None, pending=None):
        self.positions = np.asarray(positions)
        self.velocities = np.asarray(velocities)
        self.step = int(step)
        self.seed = int(seed)
        self.walker_offset = int(walker_offset)
        self.loop_step = int(loop_step)
        self.last_update_step = last_update_step
        self.bias_state = bias_state
        self.pending = pending

    def getPositions(self, asNumpy=True):
        return self.positions

    def getVelocities(self, asNumpy=True):
        return self.velocities


class SingleParticleSimulation:
    """Langevin dynamics of independent particles on a 2D potential energy surface with an optional
    VES bias (``ves/langevin_dynamics.py:15-235``)."""

    def __init__(self,
                 potential,
                 mass: int = 1,
                 temp: float = 300,
                 friction: float = 20,
                 timestep: float = 10,
                 init_state=None,
                 init_coord: np.ndarray = np.array([0, 0, 0]).reshape((1, 3)),
                 gpu: bool = True,
                 cpu_threads: int = None,
                 seed: int = None,
                 traj_in_mem: bool = True,
                 n_walkers: int = 1,
                 device: int = None,
                 precision: str = None,
                 scheme: str = "openmm_langevin",
                 dims: int = None,
                 traj_walkers: int = 1,
                 sample_every: int = None,
                 process_group=None,
                 device_update: bool = False,
                 archive_updates: bool = False,
                 communicator=None):
        # units as upstream: Da, K, 1/ps, fs
        self.mass = float(mass)
        self.temp = float(temp)
        self.friction = float(friction)
        self.timestep = float(timestep)
        self.dt_ps = self.timestep * 1e-3
        self.potential = potential
        self.init_state = init_state
        self.process_group = process_group
        self.rank, self.world = dist_info(process_group)
        self.n_walkers_total = int(n_walkers)
        lo, hi = shard(self.n_walkers_total, self.rank, self.world)
        if hi <= lo:
            raise ValueError("n_walkers must be at least the number of ranks")
        self.walker_offset, self.n_walkers = lo, hi - lo
        self.dims = int(dims) if dims is not None else (3 if self.n_walkers_total == 1 else 2)
        self.precision = precision or ("fp64" if self.n_walkers_total == 1 else "fp32")
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", 0)) if self.world > 1 else torch.cuda.current_device()
        self.device = device
        # the bias objects create their evaluator handles and CUDA tensors on the CURRENT device: make it this rank's
        torch.cuda.set_device(self.device)
        self.seed = int.from_bytes(os.urandom(7), "little") if seed is None else int(seed)
        if init_state is not None:
            self.seed = getattr(init_state, "seed", self.seed) if seed is None else self.seed
        print("Running simulation of {} walker(s) on GPU {} ({}, {} simulated dimensions).".format(
            self.n_walkers, self.device, self.precision, self.dims))

        self.ens = _lib.Ensemble(self.n_walkers, dims=self.dims, seed=self.seed, precision=self.precision, device=self.device)
        self.ens.set_integrator(self.mass, KB * self.temp, self.friction, self.dt_ps, scheme)
        self.ens.set_potential(potential.kind, potential.params)
        self.ens.set_walker_offset(self.walker_offset)

        if init_state is None:
            pos = np.asarray(init_coord, dtype=np.float64)
            if pos.ndim == 1:
                pos = pos.reshape(1, -1)
            if pos.shape[0] == 1:
                pos = np.repeat(pos, self.n_walkers, axis=0)
            elif pos.shape[0] == self.n_walkers_total and self.world > 1:
                pos = pos[lo:hi]
            if pos.shape[0] != self.n_walkers:
                raise ValueError("init_coord must have 1 or n_walkers rows")
            self._set_positions(pos)
            self.ens.init_velocities()
        else:
            self._set_positions(np.asarray(init_state.getPositions()))
            vel = np.asarray(init_state.getVelocities())
            self.ens.set_state(None, np.ascontiguousarray(self._cols(vel).T))
            self.ens.step_counter = getattr(init_state, "step", 0)

        self.traj_in_mem = traj_in_mem
        self.traj_walkers = max(1, min(int(traj_walkers), self.n_walkers))
        self.sample_every = sample_every
        self.device_update = bool(device_update)
        self.archive_updates = bool(archive_updates)       # device_update: also write model_loc + ".iter{n}" per update (costs a sync)
        self.communicator = communicator    # a pyves_b200._lib.Communicator: the device-resident loop reduces through the C ABI
        self._updater = None
        self.biased = False
        self._g0 = 0                    # loop index this run continues from (set by init_ves when the state carries a bias)
        self._last_update_step = None

    # ------------------------------------------------------------------------------------------------
    def _cols(self, a):
        """[n, C] -> [n, dims] (pad z with zeros / drop it)."""
        a = np.asarray(a, dtype=np.float64).reshape(self.n_walkers, -1)
        if a.shape[1] >= self.dims:
            return a[:, :self.dims]
        return np.concatenate([a, np.zeros((a.shape[0], self.dims - a.shape[1]))], axis=1)

    def _set_positions(self, pos):
        self.ens.set_state(np.ascontiguousarray(self._cols(pos).T), None)

    def place_walkers(self, ntrials=100, xmin=0, xmax=1, ymin=0, ymax=1):
        """Monte-Carlo placement of every walker (rule of ``ves/config_creation.py:56-68``) + fresh velocities."""
        from .config_creation import ensemble_init_coord
        ensemble_init_coord(self.ens, self.temp, ntrials, xmin, xmax, ymin, ymax)
        self.ens.init_velocities()

    def get_state(self, loop_step=0):
        pos, vel = self.ens.get_state(host=True)
        bias_state = pending = None
        if self.biased and not self.static and hasattr(self.bias, "state_dict"):
            if self._updater is not None:
                self._updater.sync_to_host()                          # device-resident optimiser state -> bias object
            bias_state = self.bias.state_dict()
            if getattr(self, "_acc", None) is not None:
                pending = dict(acc=[None if t is None else t.detach().cpu().numpy().copy() for t in self._acc],
                               sampled_at=getattr(self, "_sampled_at", None))
        return State(pos.T.copy(), vel.T.copy(), self.ens.step_counter, self.seed, self.walker_offset, loop_step=loop_step,
                     last_update_step=self._last_update_step, bias_state=bias_state, pending=pending)

    # ------------------------------------------------------------------------------------------------
    def init_ves(self, bias: Bias, static: bool = False, startafter: int = 2, learnevery: int = 50):
        """Initialize simulation with VES bias (``ves/langevin_dynamics.py:85-107``)."""
        self.biased = True
        if static:
            self.static = True
            self.startafter = None
            self.learnevery = None
        else:
            self.static = False
            if startafter < 2:
                raise ValueError("Cannot start VES updates before timestep 2.")
            self.startafter = startafter
            self.learnevery = learnevery
        self.bias = bias
        self.update_on = False
        self._updater = None
        self._g0, self._last_update_step, self._pending0 = 0, None, None
        st = getattr(self, "init_state", None)
        if not static and getattr(st, "bias_state", None) is not None and hasattr(bias, "load_state_dict"):
            # continuation of a checkpointed dynamic run: restore the update state and keep the update cadence on the
            # global loop index (the reference restarts both, ves/langevin_dynamics.py:75-76 resumes positions only)
            bias.load_state_dict(st.bias_state)
            self._g0 = int(st.loop_step)
            self._last_update_step = st.last_update_step
            self._pending0 = st.pending

    # ------------------------------------------------------------------------------------------------
    def _attach_ensemble_bias(self):
        """The bias starts acting (ves/langevin_dynamics.py:141-145): defines the basis for the moment kernels."""
        if self._attached:
            return
        if self.device_update and self._updater is None and DeviceUpdater.supports(self.bias):
            self._updater = DeviceUpdater(self.bias, self.ens)    # optimiser state + coefficients stay on the GPU
        else:
            self.bias.force.apply(self.ens)
        self._attached = True

    def _archive_device_update(self):
        if self.archive_updates and self.bias.model_loc is not None:
            self._updater.sync_to_host()
            self.bias._export(self.bias.model_loc, self.bias.model_loc + ".iter{}".format(self.bias.n_updates))

    def _update_bias(self, i, ves_update_params):
        self._last_update_step = i
        if isinstance(self.bias, VESBias_Ensemble):
            self._attach_ensemble_bias()
            if self._updater is not None and self._acc is None:
                # lean path: moments into the updater's persistent flat buffer, ONE in-place all-reduce, one optimiser
                # kernel, coefficient hand-over -- four library calls and one NCCL call per update, nothing else
                bufs = self._updater.moment_buffers(self.bias.needs_covariance)
                self.ens.moments(cov=self.bias.needs_covariance, out=bufs)
                self._updater.reduce_flat(self.process_group, comm=self.communicator)
                self._updater.update(*bufs)
                self._archive_device_update()
                return
            self._sample_moments()
            sw, sf, sff = reduce_moments(*self._acc, group=self.process_group)   # one allreduce per update
            self._acc = None
            if self._updater is not None:
                self._updater.update(sw, sf, sff)                         # enqueue only: no host round trip
                self._archive_device_update()
                return
            self.bias.update_from_moments(sw, sf, sff)
        else:
            self.bias.update(self.traj, **ves_update_params)
        self.bias.force.apply(self.ens)

    def _sample_moments(self):
        """Accumulate sum_w, sum_w f, (sum_w f f) of the current walkers for the next ensemble update."""
        if getattr(self, "_sampled_at", None) == self.ens.step_counter and self._acc is not None:
            return
        s = self.ens.moments(cov=self.bias.needs_covariance)
        if self._acc is None:
            self._acc = list(s)
        else:
            for k in range(3):
                if s[k] is not None:
                    self._acc[k] += s[k]
        self._sampled_at = self.ens.step_counter

    def _events(self, i, nsteps, chkevery, trajevery, energyevery):
        """Smallest step index > i at which the host must do something."""
        nxt = nsteps
        for every in (chkevery, energyevery):
            if every:
                nxt = min(nxt, (i // every + 1) * every)
        if self.biased and not self.static:
            g0 = self._g0
            g = g0 + i                                   # update cadence runs on the global loop index
            if g < self.startafter:
                nxt = min(nxt, self.startafter - g0)
            nxt = min(nxt, max((g // self.learnevery + 1) * self.learnevery, self.startafter + 1) - g0)
            if isinstance(self.bias, VESBias_Ensemble) and self.sample_every:
                nxt = min(nxt, (g // self.sample_every + 1) * self.sample_every - g0)
        return nxt

    def __call__(self,
                 nsteps: int = 1000,
                 chkevery: int = 500,
                 trajevery: int = 1,
                 energyevery: int = 1,
                 chkfile="./chk_state.pkl",
                 trajfile="./traj.dat",
                 energyfile="./energies.dat",
                 ves_update_params={}):
        self.traj = None
        self.PE = []
        self.KE = []
        self._acc = None
        self._attached = False
        g0 = self._g0
        if self.biased and not self.static and g0 > 0 and isinstance(self.bias, VESBias_Ensemble):
            # continuation: the bias was already acting when the checkpoint was written
            if g0 > self.startafter or self._last_update_step == g0:
                self._attach_ensemble_bias()
            if self._pending0 is not None:
                dev = torch.device("cuda", self.device)
                self._acc = [None if a is None else torch.as_tensor(a, device=dev) for a in self._pending0["acc"]]
                self._sampled_at = self._pending0["sampled_at"]
        self._traj_written = 0
        self.traj_ensemble = None
        frames = []
        i = 0
        i_last = 0
        ens = self.ens
        write = self.rank == 0
        ftraj = open(trajfile, "w") if (write and trajfile and trajevery) else None
        fener = open(energyfile, "w") if (write and energyfile and energyevery) else None
        self._ftraj, self._trajevery = ftraj, trajevery
        if ftraj:
            ftraj.write("# t[ps]    x [nm]    y [nm]    z[nm]\n")
        if fener:
            fener.write("# t[ps]    PE [kJ/mol]    KE [kJ/mol]\n")
        try:
            while i < nsteps:
                # ---- VES bookkeeping at iteration i (ves/langevin_dynamics.py:134-169)
                if self.biased:
                    if not self.static:
                        g = g0 + i
                        if g == self._last_update_step and i == 0 and g0 > 0:
                            pass                                    # the checkpoint was written after the update at this index
                        elif g == self.startafter:
                            print("[VES] {} bias: Initializing dynamic bias at timestep {}.".format(type(self.bias).__name__, g))
                            self._sync_traj(frames)
                            self._update_bias(g, ves_update_params)
                        elif g > self.startafter and g % self.learnevery == 0:
                            print("[VES] {} bias: Updating dynamic bias at timestep {}.".format(type(self.bias).__name__, g))
                            self._sync_traj(frames)
                            self._update_bias(g, ves_update_params)
                        elif (isinstance(self.bias, VESBias_Ensemble) and self.sample_every and g > self.startafter
                              and g % self.sample_every == 0):
                            self._sample_moments()
                    elif i == 0:
                        print("[VES] {} bias: Initializing static bias.".format(type(self.bias).__name__))
                        self.bias.force.apply(ens)
                if i > 0 and chkevery and i % chkevery == 0:
                    self._dump_state(chkfile, i, loop_step=g0 + i)
                if energyevery and i % energyevery == 0:
                    PE, KE = ens.energies(host=True)
                    self.PE.append(float(PE[0]) if self.n_walkers_total == 1 else float(np.mean(PE)))
                    self.KE.append(float(KE[0]) if self.n_walkers_total == 1 else float(np.mean(KE)))
                    if fener:
                        fener.write("{:10.5f}\t{:10.7f}\t{:10.7f}\n".format(i * self.dt_ps, self.PE[-1], self.KE[-1]))
                # ---- advance to the next event, recording frames inside the kernel
                nxt = self._events(i, nsteps, chkevery, trajevery, energyevery)
                if trajevery:
                    head = (-i) % trajevery                     # steps until the next frame index
                    if head and head < nxt - i:
                        ens.step(head)
                        i += head
                    if i % trajevery == 0 and i < nxt:
                        t = ens.step_record(nxt - i, every=trajevery, n_rec=self.traj_walkers)
                        frames.append((i, t))
                        i = nxt
                    elif i < nxt:
                        ens.step(nxt - i)
                        i = nxt
                else:
                    ens.step(nxt - i)
                    i = nxt
                i_last = i
                if len(frames) >= 64:
                    self._sync_traj(frames)
            self._sync_traj(frames)
            # Finalize: checkpoint named after the last loop index, as upstream (:195); written while the device-resident
            # update state is still alive so that it carries the optimiser
            self._dump_state(chkfile, max(i_last - 1, 0), loop_step=g0 + i_last)
        finally:
            if getattr(self, "_updater", None) is not None:       # coefficients, optimiser state, target back to the bias object
                self._updater.sync_to_host()
                self._updater.close()
                self._updater = None
                self.bias.force.apply(ens)
                self.bias._export(self.bias.model_loc)
            self._ftraj = None
            if ftraj:
                ftraj.close()
            if fener:
                fener.close()

    # ------------------------------------------------------------------------------------------------
    def _sync_traj(self, frames):
        """Move recorded frames to the host: ``self.traj`` ([T, 3] of walker 0, plus
        ``self.traj_ensemble`` [T, traj_walkers, dims]) and the trajectory file."""
        if not frames:
            return
        ftraj, trajevery = self._ftraj, self._trajevery
        chunks = []
        for i0, t in frames:
            chunks.append(t.permute(0, 2, 1).cpu().numpy().astype(np.float64))      # [frames, rec, dims]
        frames.clear()
        new = np.concatenate(chunks, axis=0)
        if self.traj_in_mem:
            first = new[:, 0, :]
            if first.shape[1] < 3:
                first = np.concatenate([first, np.zeros((first.shape[0], 3 - first.shape[1]))], axis=1)
            self.traj = first if self.traj is None else np.vstack((self.traj, first))
            if self.traj_walkers > 1:
                prev = self.traj_ensemble
                self.traj_ensemble = new if prev is None else np.concatenate([prev, new], axis=0)
        if ftraj is not None:
            k0 = self._traj_written
            for k in range(new.shape[0]):
                p = new[k, 0]
                z = p[2] if p.shape[0] > 2 else 0.0
                ftraj.write("{:10.5f}\t{:10.7f}\t{:10.7f}\t{:10.7f}\n".format((k0 + k) * trajevery * self.dt_ps, p[0], p[1], z))
        self._traj_written += new.shape[0]

    def _dump_state(self, ofilename, i, loop_step=None):
        t = i * self.dt_ps
        print("Checkpoint at {:10.7f} ps".format(t))
        if not ofilename:
            return
        name = ofilename if self.world == 1 else "{}.rank{}".format(ofilename, self.rank)
        with open(name, "wb") as fh:
            pickle.dump(self.get_state(loop_step=i if loop_step is None else loop_step), fh)
