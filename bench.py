#!/usr/bin/env python
"""bench.py -- walker-steps/s of the biased-Langevin VES hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): 2D double well, 16 x 16 tensor-product Legendre bias on
[-2.5, 2.5] x [-2, 2], well-tempered target (bias factor 10, 100 x 100 grid, refreshed every 10
updates), averaged-SGD coefficient update every 500 MD steps, 10^6 walkers per GPU (weak scaling;
walkers shard by global id, one NCCL allreduce of the 257 basis moments per update).

One bench "step" = one VES iteration = 500 fused Langevin steps of every walker + the moment
reduction + the coefficient update + the bias re-upload.  ``value`` is measured with the walker state
resident in HBM; ``e2e`` drives the same iteration through the C ABI with HOST buffers (state uploaded
from pinned memory and read back every iteration).  ``--impl reference`` times the CPU
implementation of the same path (the oracle's OpenMP C port: the reference itself is Python + OpenMM,
and OpenMM is absent from this image) on the host cores, on the same workload: in-box averages, well-tempered refresh
every 10 updates, preconditioned averaged SGD.

Besides the headline the JSON line carries ``configs``: short legs of the other BASELINE.json configurations run in
the same process (cfg1 static LegendreBasis1D(5), cfg3 slip bond + second-order update at 10^7 walkers in total
= strong scaling over the ranks, cfg4 NN bias with the device-resident Adam loop, cfg5 K x K bases for K = 8 / 32 /
64 and a 64 x 64 strong-scaling point), each with ``value``, ``kernel``, ``kernel_ms`` and ``roofline.frac``; and
``comm_ms``, the time of the one all-reduce per update (CUDA events, max over ranks).  ``--scaling strong`` makes
the HEADLINE itself a strong-scaling run (``--walkers-total`` walkers sharded over the ranks).
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

KB = 8.31446261815324e-3
KB_PY = 8.3145e-3
TEMP = 300.0
MINMAX = (-2.5, 2.5, -2.0, 2.0)
FLOPS_PER_WALKER_STEP = {8: 4 * 64 + 4 * 8 + 6 * 14 + 32, 16: 1300, 32: 4628, 64: 17428}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--walkers", type=int, default=1000000, help="walkers per GPU")
    ap.add_argument("--stride", type=int, default=500, help="MD steps per VES iteration")
    ap.add_argument("--basis", type=int, default=16, help="Legendre functions per axis")
    ap.add_argument("--mu", type=float, default=0.5, help="averaged-SGD step size (kJ/mol)")
    ap.add_argument("--fes-updates", type=int, default=600,
                    help="untimed VES updates (100 MD steps each) run after the timed region before the FES RMSE is taken; 0 = skip")
    ap.add_argument("--seed", type=int, default=20221019)
    ap.add_argument("--update", default="device", choices=["device", "host"],
                    help="coefficient update of the resident-state loop: 'device' = optimiser, target refresh and coefficient "
                         "hand-over on the GPU (DeviceUpdater), 'host' = numpy optimiser + re-upload (VESBias_Ensemble)")
    ap.add_argument("--e2e-chunks", type=int, default=16,
                    help="e2e with --update device: the ensemble is cut into about this many handles on their own streams (at multiples "
                         "of a wave of the step kernel, pyves_step_wave_walkers) so that the host<->device copies of one chunk overlap the "
                         "step kernel of another (1 = one handle, serial copies).  Measured at 1 GPU with the 4.8 ms step kernel: 3 / 5 / 7 / 14 "
                         "chunks 9.29 / 9.28 / 9.24 / 9.47e10; at 8 GPUs more chunks help more (the exposed head and tail copies cost more "
                         "when eight processes share the host)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--comm", default="native", choices=["native", "torch"],
                    help="the all-reduce of the moments: 'native' = pyves_allreduce_moments (NCCL bound by the C-ABI library itself, "
                         "communicator bootstrapped through the torch process group), 'torch' = torch.distributed.all_reduce")
    ap.add_argument("--step-kernel", default="auto", choices=["auto", "cuda", "tensor"],
                    help="2D Legendre step kernel of the headline: 'auto' = the library's table (tensor-core contraction for FP32 bases other than the static 8 x 8 "
                         "double-well kernel), 'cuda' = CUDA-core column sweep (pyves_set_tuning 6), 'tensor' = tensor cores forced (7)")
    ap.add_argument("--no-configs", action="store_true", help="skip the legs of the other BASELINE.json configurations")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --walkers per GPU (default); strong: --walkers-total sharded over the ranks")
    ap.add_argument("--walkers-total", type=int, default=10000000, help="ensemble size of --scaling strong")
    ap.add_argument("--ref-budget", type=float, default=280.0, help="--impl reference: wall-clock budget (s) of the whole run")
    a = ap.parse_args()
    if a.scaling == "strong":
        world = int(os.environ.get("WORLD_SIZE", 1))
        a.walkers = a.walkers_total // world
    return a


# ---------------------------------------------------------------------------------------------------
class ClockMonitor(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU while the timed region runs (NVML)."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, index, period=0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.power = [], set(), []
        self.stop_flag = threading.Event()
        self.max_mhz = None
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        while not self.stop_flag.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                self.power.append(self.nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def result(self):
        self.stop_flag.set()
        if self.is_alive():
            self.join()
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": "NVML unavailable"}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "power_w_max": max(self.power) if self.power else None, "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------------
def make_bias(args, beta):
    from pyves_b200.basis import LegendreBasis2D
    from pyves_b200.bias import VESBias_Ensemble
    from pyves_b200.target import Target_WellTempered_2D
    d = args.basis - 1
    basis = LegendreBasis2D(d, d, *MINMAX)                      # w = 0 at start
    target = Target_WellTempered_2D(100, 100, bias_factor=10.0)
    return VESBias_Ensemble(basis, target, beta, "averaged_sgd",
                            {"mu": args.mu, "averaging": "ewma", "decay": 0.8, "precondition": 1.0},
                            model_loc=None, target_update_every=10)


def make_ensemble(args, rank, device):
    from pyves_b200 import _lib
    ens = _lib.Ensemble(args.walkers, dims=2, seed=args.seed, precision="fp32", device=device)
    ens.set_integrator(1.0, KB * TEMP, 20.0, 0.01)
    ens.set_potential(_lib.POT_DOUBLE_WELL_2D)
    ens.set_walker_offset(rank * args.walkers)
    failed = ens.init_positions_mc(MINMAX, 1.0 / (KB_PY * TEMP))
    assert failed == 0
    ens.init_velocities()
    return ens


def peak_all(_lib, local, rank, world, dev):
    """FP32 FMA peak of rank 0's GPU for the legs' roofline (measured once, before the timed region; see run_ours)."""
    return _PEAK.get("tflops")


_PEAK = {}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from pyves_b200 import _lib
    from pyves_b200.parallel import reduce_moments
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_cores = None
    if world > 1 and os.environ.get("PYVES_NO_NUMA_BIND", "0") != "1":
        from pyves_b200.parallel import bind_to_gpu_numa_node
        numa_cores = bind_to_gpu_numa_node(local)          # before any pinned allocation: host buffers land next to this GPU
    beta = 1.0 / (KB_PY * TEMP)
    from pyves_b200.potentials import _Potential
    _Potential.quiet = True
    bias = make_bias(args, beta)
    ens = make_ensemble(args, rank, local)
    ens.set_tuning({"auto": -1, "cuda": 6, "tensor": 7}[args.step_kernel])
    bias.force.apply(ens)
    K = args.basis * args.basis
    n_total = args.walkers * world
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    step_events, comm_events = [], []

    def barrier():
        if world > 1:
            dist.barrier()

    from pyves_b200.bias import DeviceUpdater
    upd = DeviceUpdater(bias, ens) if args.update == "device" else None
    comm = _lib.Communicator.from_process_group(local) if (world > 1 and args.comm == "native") else None
    _PEAK["comm"] = comm

    def iteration(timed):
        flush.zero_()                                            # evict the 126 MB L2 between iterations
        if timed:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
        ens.step(args.stride)
        if timed:
            b.record()
            step_events.append((a, b))
        if upd is not None:
            bufs = upd.moment_buffers(False)
            ens.moments(out=bufs)                                # 1 + K doubles into the updater's flat buffer
            if timed and world > 1:
                c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                c0.record()
                upd.reduce_flat(comm=comm)                       # one NCCL all-reduce, in place
                c1.record()
                comm_events.append((c0, c1))
            else:
                upd.reduce_flat(comm=comm)
            upd.update(*bufs)                                    # optimiser kernel + target + coefficient hand-over, all enqueued
        else:
            sw, sf, _ = reduce_moments(*ens.moments())           # one NCCL allreduce of 1 + K doubles
            bias.update_from_moments(sw, sf)                     # host optimiser on K coefficients
            bias.force.apply(ens)                                # re-upload the coefficients

    peak = _lib.fp32_fma_peak(local) if rank == 0 else None     # before the timed region
    _PEAK["tflops"] = peak
    for _ in range(args.warmup):
        iteration(False)
    torch.cuda.synchronize()
    barrier()
    mon = ClockMonitor(local, period=float(os.environ.get("PYVES_CLOCK_PERIOD", "0.02")))
    mon.start()
    l0 = _lib.launch_count()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        iteration(True)
    t1.record()
    torch.cuda.synchronize()
    barrier()
    clocks = mon.result()
    launches = _lib.launch_count() - l0
    elapsed = torch.tensor([t0.elapsed_time(t1) * 1e-3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(elapsed, op=dist.ReduceOp.MAX)
    elapsed = float(elapsed)
    value = n_total * args.stride * args.steps / elapsed
    kern_ms = float(np.mean([a.elapsed_time(b) for a, b in step_events]))
    # the one collective per update: CUDA events around the all-reduce on the launching stream (includes the wait for the
    # slowest rank's moments, i.e. the skew), mean per iteration, MAX over ranks
    comm_t = torch.tensor([float(np.mean([a.elapsed_time(b) for a, b in comm_events])) if comm_events else 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(comm_t, op=dist.ReduceOp.MAX)
    comm_ms = float(comm_t)

    # ---- end to end through the C ABI with host buffers -------------------------------------------
    # every iteration: walker state from pinned host memory -> GPU, the VES iteration, then the result (the 1 + K
    # basis moments) and the new walker state back to the host.  With --update device the optimiser itself stays on
    # the GPU (the coefficients are not a per-step input); with --update host it is the numpy optimiser + re-upload.
    e2e = None
    if not args.no_e2e:
        hpos = torch.empty((2, args.walkers), dtype=torch.float32).pin_memory().numpy()
        hvel = torch.empty((2, args.walkers), dtype=torch.float32).pin_memory().numpy()
        hmom = torch.empty(K + 1, dtype=torch.float64).pin_memory()
        ens.get_state(out=(hpos, hvel))

        # chunk boundaries: the tensor-core step kernel keeps one large CTA per SM, so a launch costs whole waves of
        # wave_walkers() walkers; chunks are cut at multiples of a wave (a partial wave only in the last chunk)
        bounds = [0, args.walkers]
        if upd is not None and args.e2e_chunks > 1:
            wave = ens.wave_walkers()
            if wave > 0:
                waves_total = -(-args.walkers // wave)
                per_chunk = max(1, int(round(waves_total / args.e2e_chunks))) * wave
            else:
                per_chunk = -(-args.walkers // args.e2e_chunks)
            bounds = list(range(0, args.walkers, per_chunk)) + [args.walkers]
        nch = len(bounds) - 1
        if nch > 1:
            # chunked pipeline: walkers are independent between updates, so the ensemble is split into nch handles
            # (global walker ids preserved), each on its own stream; all copies are asynchronous from pinned memory
            subs, streams, hp, hv = [], [torch.cuda.Stream(device=dev) for _ in range(nch)], [], []
            for c in range(nch):
                c0, c1 = bounds[c], bounds[c + 1]
                e = _lib.Ensemble(c1 - c0, dims=2, seed=args.seed, precision="fp32", device=local)
                e.set_integrator(1.0, KB * TEMP, 20.0, 0.01)
                e.set_potential(_lib.POT_DOUBLE_WELL_2D)
                e.set_walker_offset(rank * args.walkers + c0)
                e.set_tuning({"auto": -1, "cuda": 6, "tensor": 7}[args.step_kernel])
                e.step_counter = ens.step_counter
                upd.add_handle(e)
                subs.append(e)
                hp.append(torch.from_numpy(np.ascontiguousarray(hpos[:, c0:c1])).pin_memory().numpy())
                hv.append(torch.from_numpy(np.ascontiguousarray(hvel[:, c0:c1])).pin_memory().numpy())
            # per-chunk moment buffers allocated ONCE on the main stream and written through out=: no tensor allocated on a
            # side stream is consumed on the main stream (the caching allocator could hand such a block out again too early)
            mom = [(torch.zeros(1, dtype=torch.float64, device=dev), torch.zeros(K, dtype=torch.float64, device=dev), None) for _ in range(nch)]

        def e2e_chunked():
            main = torch.cuda.current_stream(dev)
            for c in range(nch):
                streams[c].wait_stream(main)                     # the coefficient hand-over was enqueued on the main stream
                with torch.cuda.stream(streams[c]):
                    # pyves_iterate_pinned: H2D of this chunk's walkers from pinned memory, the steps, the chunk's moments, D2H of
                    # the walkers back into the pinned buffers -- one C call, all enqueued on this chunk's stream
                    subs[c].iterate_pinned(hp[c], hv[c], args.stride, mom[c])
            for c in range(nch):
                main.wait_stream(streams[c])
            sw = sum(m[0] for m in mom)
            sf = sum(m[1] for m in mom)
            sw, sf, _ = reduce_moments(sw, sf)
            upd.update(sw, sf)
            hmom.copy_(torch.cat([sw.reshape(-1), sf.reshape(-1)]), non_blocking=True)       # D2H of the result (moments)
            torch.cuda.synchronize(dev)                          # walker state and moments are on the host

        def e2e_iteration():
            if nch > 1:
                return e2e_chunked()
            ens.set_state(hpos, hvel)                            # H2D of this iteration's inputs
            ens.step(args.stride)
            if upd is not None:
                sw, sf, _ = reduce_moments(*ens.moments())
                upd.update(sw, sf)
                hmom.copy_(torch.cat([sw.reshape(-1), sf.reshape(-1)]), non_blocking=True)   # D2H of the result (moments)
                ens.get_state(out=(hpos, hvel))                  # D2H of the walker state (synchronises)
                return
            sw, sf, _ = ens.moments(host=True)                   # D2H of the result (moments)
            ens.get_state(out=(hpos, hvel))                      # D2H of the walker state
            if world > 1:
                sw, sf, _ = reduce_moments(torch.from_numpy(sw).to(dev), torch.from_numpy(sf).to(dev))
            bias.update_from_moments(sw, sf)
            bias.force.apply(ens)

        e2e_iteration()
        torch.cuda.synchronize()
        barrier()
        w0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_iteration()
        torch.cuda.synchronize()
        barrier()
        w = torch.tensor([time.perf_counter() - w0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(w, op=dist.ReduceOp.MAX)
        state_bytes = 2 * 2 * args.walkers * 4
        e2e = {"value": n_total * args.stride * args.steps / float(w), "unit": "walker-steps/s",
               "h2d_bytes_per_step": state_bytes + (0 if upd is not None else K * 8), "d2h_bytes_per_step": state_bytes + (K + 1) * 8,
               "chunks": nch, "chunk_walkers": [bounds[c + 1] - bounds[c] for c in range(nch)], "numa_bound_cores": None if numa_cores is None else len(numa_cores)}
        if nch > 1:
            for e in subs:
                upd.handles.remove(e)
                e.close()

    if upd is not None:                                          # hand the state back to the host-side objects
        upd.sync_to_host()
        upd.close()
        bias.force.apply(ens)

    # ---- the other BASELINE.json configurations, short legs (all ranks take part)
    legs = None
    if not args.no_configs:
        legs = config_legs(args, rank, world, local, dev, peak_all(_lib, local, rank, world, dev))
        if rank == 0 and world == 1:
            legs["cfg0_single_walker"] = single_walker_rate(local, args.seed)

    # ---- accuracy: continue the optimisation (untimed) and compare the FES with the analytic surface
    fes_out = None
    if args.fes_updates > 0:
        for _ in range(args.fes_updates):
            ens.step(100)
            sw, sf, _ = reduce_moments(*ens.moments())
            bias.update_from_moments(sw, sf)
            bias.force.apply(ens)
        if rank == 0:
            fes_out = fes_rmse(bias, args)
            fes_out["md_steps_total"] = (args.warmup + args.steps * (2 if not args.no_e2e else 1) + 1) * args.stride + 100 * args.fes_updates

    if rank == 0:
        flops = FLOPS_PER_WALKER_STEP.get(args.basis)
        roof = None
        if flops:
            achieved = args.walkers * args.stride * flops / (kern_ms * 1e-3) * 1e-12
            tensor_path = args.step_kernel == "tensor" or (args.step_kernel == "auto" and 14 <= args.basis <= 64)
            roof = {"bound": "fp32-fma (CUDA cores; neither HBM nor tensor bound)", "kernel": "langevin2d_kernel",
                    "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                    "peak_source": "FFMA micro-benchmark run in this process (pyves_fp32_fma_peak); nominal 74.4",
                    "flops_per_walker_step": flops, "kernel_ms": kern_ms,
                    "kernel_share_of_step": kern_ms * args.steps / (elapsed * 1e3),
                    **ncu_traffic(args)}
            if tensor_path:
                roof.update(tensor_roofline(args.basis, args.walkers * args.stride / (kern_ms * 1e-3)))
            # the HBM view of the same kernel, against the driver-measured copy bandwidth: why "hbm" is not the bound
            hbm_peak, hbm_src = 6650.0, "fallback of B200_PROFILING.md"
            try:
                with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "MEASURED_PEAKS.json")) as fh:
                    hbm_peak, hbm_src = float(json.load(fh)["hbm_gbs"]), "MEASURED_PEAKS.json"
            except (OSError, KeyError, ValueError):
                pass
            alg_bytes = 32.0 * args.walkers                       # state read + written once per launch, 16 B each way per walker
            hbm_gbs = alg_bytes / (kern_ms * 1e-3) * 1e-9
            roof["hbm"] = {"algorithmic_bytes_per_launch": alg_bytes, "achieved": hbm_gbs, "peak": hbm_peak, "unit": "GB/s",
                           "frac": hbm_gbs / hbm_peak, "peak_source": hbm_src}
        out = {"metric": "walker-steps/s (2D double well, Legendre bias)", "value": value, "unit": "walker-steps/s",
               "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed / args.steps * 1e3,
               "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
               "config": {"workload": workload_string(args),
                          "walkers_total": n_total, "md_steps_per_bench_step": args.stride, "parallelism": "walkers sharded, dp%d" % world,
                          "l2": "flushed between steps (256 MiB memset); state 16 MB/GPU is read once per 500 MD steps",
                          "update": args.update},
               "clocks": clocks, "gpu_launches": int(launches), "roofline": roof, "comm_ms": comm_ms,
               "comm_note": "one all-reduce of %d doubles per update (%s); CUDA events around it on the launching stream (includes the wait "
                            "for the slowest rank), mean per iteration, max over ranks"
                            % (K + 1, "pyves_allreduce_moments, C ABI" if comm is not None else "torch.distributed")}
        if legs:
            out["configs"] = legs
        if e2e:
            out["e2e"] = e2e
        if fes_out:
            out["fes"] = fes_out
        if world == 1 and not args.no_cpu:
            out["cpu_baseline"] = cpu_baseline(args, seconds=12.0)
        print(json.dumps(out), flush=True)
    if comm is not None:
        comm.close()
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------
# Legs of the other BASELINE.json configurations (same process, after the headline's timed region).  Every leg is a few
# VES iterations of that configuration through the same C ABI; `value` is whole-job walker-steps/s of the full iteration
# (MD steps + moments + all-reduce + update), `kernel_ms` the mean step-kernel launch (CUDA events), `roofline.frac` that
# kernel's algorithmic flops against the FP32 FMA peak measured in this process.
LEG_FLOPS = {"cfg1": 70, "cfg3": 133, "cfg4": 9635, 8: 404, 16: 1300, 32: 4628, 64: 17428}


def _leg_ensemble(n, gid0, pot_kind, pot_params, box, local, seed, ntrials=2000):
    from pyves_b200 import _lib
    e = _lib.Ensemble(n, dims=2, seed=seed, precision="fp32", device=local)
    e.set_integrator(1.0, KB * TEMP, 20.0, 0.01)
    e.set_potential(pot_kind, pot_params)
    e.set_walker_offset(gid0)
    assert e.init_positions_mc(box, 1.0 / (KB_PY * TEMP), ntrials=ntrials) == 0
    e.init_velocities()
    return e


def _leg_time(torch, dist, world, dev, ens, stride, iters, warm, update):
    """(seconds per iteration MAX over ranks, mean step-kernel ms) of `iters` iterations: ens.step(stride); update()."""
    for _ in range(warm):
        ens.step(stride)
        update()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ev = []
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ens.step(stride)
        b.record()
        ev.append((a, b))
        update()
    t1.record()
    torch.cuda.synchronize()
    el = torch.tensor([t0.elapsed_time(t1) * 1e-3 / iters], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(el, op=dist.ReduceOp.MAX)
    return float(el), float(np.mean([a.elapsed_time(b) for a, b in ev]))


def config_legs(args, rank, world, local, dev, peak):
    import torch
    import torch.distributed as dist
    from pyves_b200 import _lib, basis, bias, target
    from pyves_b200.bias import DeviceUpdater
    from pyves_b200.parallel import shard
    beta = 1.0 / (KB_PY * TEMP)
    out = {}

    def row(name, workload, n_rank, n_total, stride, sec, kern_ms, kernel, flops, scaling, extra=None):
        r = {"workload": workload, "walkers_total": int(n_total), "md_steps_per_iteration": stride, "scaling": scaling,
             "value": n_total * stride / sec, "unit": "walker-steps/s", "ms_per_iteration": sec * 1e3, "kernel": kernel, "kernel_ms": kern_ms}
        if peak and flops:
            ach = n_rank * stride * flops / (kern_ms * 1e-3) * 1e-12
            r["roofline"] = {"bound": "fp32-fma", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "flops_per_walker_step": flops}
        if extra:
            r.update(extra)
        out[name] = r

    def updater_leg(name, workload, make_bias, pot_kind, pot_params, box, n_rank, gid0, n_total, stride, iters, kernel, flops, scaling, cov=False,
                    variant=-1):
        b = make_bias()
        e = _leg_ensemble(n_rank, gid0, pot_kind, pot_params, box, local, args.seed)
        e.set_tuning(variant)
        upd = DeviceUpdater(b, e)

        def update():
            bufs = upd.moment_buffers(cov)
            e.moments(cov=cov, out=bufs)
            upd.reduce_flat(comm=_PEAK.get("comm"))
            upd.update(*bufs)
        sec, kms = _leg_time(torch, dist, world, dev, e, stride, iters, 1, update)
        extra = {"update": "device-resident (%s)" % b.optimizer_type}
        if "tc_kernel" in kernel:
            K = int(name.split("_")[1].split("x")[0]) if name.startswith("cfg5_") else args.basis
            extra["tensor"] = tensor_roofline(K, n_rank * stride / (kms * 1e-3))["tensor"]
        elif "BiasPwl1D" in kernel:
            flops_note = ("%d flop per walker-step is the algorithmic work of evaluating the network (SURVEY 8(d)); this kernel does not "
                          "evaluate it -- for a ReLU net of one input dV/ds is piecewise constant and is looked up by a binary search in "
                          "shared memory -- so the FP32 fraction is not a bound here (it exceeds 1)" % flops)
            extra["roofline_note"] = flops_note
        elif "tc_relu_kernel" in kernel:
            ex = 2.0 * 2.0 * 96 * 48                  # 2 MMAs (mask x hi, mask x lo) x 2 N K flops per walker-step, N = 2 H, K = H
            extra["tensor"] = {"executed_flops_per_walker_step": ex, "achieved": ex * n_rank * stride / (kms * 1e-3) * 1e-12, "unit": "TFLOP/s"}
        elif "tc_mlp_kernel" in kernel:
            ex = 12.0 * 48 * 48
            extra["tensor"] = {"executed_flops_per_walker_step": ex, "achieved": ex * n_rank * stride / (kms * 1e-3) * 1e-12, "unit": "TFLOP/s"}
        row(name, workload, n_rank, n_total, stride, sec, kms, kernel, flops, scaling, extra)
        upd.close()
        e.close()

    # cfg1 (configs[0] at ensemble size): static LegendreBasis1D(5) along x, no updates, 2^20 walkers per GPU
    n1 = 1 << 20
    e = _leg_ensemble(n1, rank * n1, _lib.POT_DOUBLE_WELL_2D, (), MINMAX, local, args.seed)
    bias.StaticBias_SingleParticle(basis.LegendreBasis1D(5, -2.5, 2.5, 'x', weights=np.array([-11.8, 4.2, 13.1, -2.5, -9.0, 0.4])), None).force.apply(e)
    sec, kms = _leg_time(torch, dist, world, dev, e, 1000, 3, 1, lambda: None)
    row("cfg1", "configs[0] shape: 2D double well + static LegendreBasis1D(5, -2.5, 2.5, 'x'), 2^20 walkers per GPU", n1, n1 * world, 1000, sec, kms,
        "langevin2d_pair_kernel<DoubleWell, Legendre1D<5>>", LEG_FLOPS["cfg1"], "weak")
    e.close()

    # cfg3 (configs[2]): slip bond, LegendreBasis1D(12, -3, 6), uniform target, second-order averaged SGD, 10^7 walkers in TOTAL
    lo, hi = shard(10 ** 7, rank, world)
    updater_leg("cfg3", "configs[2]: slip bond + LegendreBasis1D(12,-3,6,'x'), uniform target, 2nd-order averaged SGD (13x13 covariance), "
                "10^7 walkers sharded over the ranks",
                lambda: bias.VESBias_Ensemble(basis.LegendreBasis1D(12, -3, 6, 'x', weights=np.zeros(13)), target.Target_Uniform_HardSwitch_x(200), beta,
                                              "averaged_sgd", {"mu": 0.5, "second_order": True, "averaging": "ewma", "decay": 0.8}, target_update_every=0),
                _lib.POT_SLIP_BOND_2D, (0, 0, 1, 5, 4, 0, 2), (-8, 10, -6, 8), hi - lo, lo, 10 ** 7, args.stride, 3,
                "langevin2d_pair_kernel<SlipBond, Legendre1D<12>>", LEG_FLOPS["cfg3"], "strong", cov=True)

    # cfg4 (configs[3]): Szabo-Berezhkovskii, NNBasis1D [48, 48, 48], Adam 1e-3, running-average gradients, 10^6 walkers per GPU
    def nn_bias():
        torch.manual_seed(0)
        return bias.VESBias_Ensemble(basis.NNBasis1D(min=-4, max=4, axis='x', hidden_layer_sizes=[48, 48, 48]), target.Target_Uniform_HardSwitch_x(200),
                                     beta, "Adam", {"lr": 1e-3}, target_update_every=0, gradient_average='running')
    updater_leg("cfg4", "configs[3]: Szabo-Berezhkovskii + NNBasis1D [48,48,48] (ReLU), Adam 1e-3, running-average gradients, uniform target, "
                "10^6 walkers per GPU", nn_bias, _lib.POT_SZABO_BEREZHKOVSKII, (2.2, 4.0, 4.04, 4.84), (-4, 4, -4, 4), 1000000, rank * 1000000,
                1000000 * world, args.stride, 2, "langevin2d_kernel<SzaboBerezhkovskii, BiasPwl1D> (ReLU net compiled to its piecewise-constant dV/ds: "
                "binary search in shared memory; parameter-gradient moments through the pieces)", LEG_FLOPS["cfg4"], "weak")
    updater_leg("cfg4_tensor_cores", "configs[3] with the network itself evaluated every step, 48 x 48 layer on the tensor cores (pyves_set_tuning 7)", nn_bias,
                _lib.POT_SZABO_BEREZHKOVSKII, (2.2, 4.0, 4.04, 4.84), (-4, 4, -4, 4), 1000000, rank * 1000000, 1000000 * world, args.stride, 2,
                "langevin2d_tc_relu_kernel<48, 5> (0/1 mask operand against W2 diag(u) over W2 diag(c))", LEG_FLOPS["cfg4"], "weak", variant=7)
    updater_leg("cfg4_cuda_cores", "configs[3] with the step kernel forced onto the CUDA cores (pyves_set_tuning 6)", nn_bias, _lib.POT_SZABO_BEREZHKOVSKII,
                (2.2, 4.0, 4.04, 4.84), (-4, 4, -4, 4), 1000000, rank * 1000000, 1000000 * world, args.stride, 2,
                "langevin2d_kernel<SzaboBerezhkovskii, Mlp<48>> (CUDA cores, FFMA2 input sweep)", LEG_FLOPS["cfg4"], "weak", variant=6)

    # cfg5 (configs[4]): K x K bases, uniform target; weak points at 10^6 walkers per GPU and a 64 x 64 strong point at 10^7 in total
    def leg2d(K):
        return lambda: bias.VESBias_Ensemble(basis.LegendreBasis2D(K - 1, K - 1, *MINMAX), target.Target_Uniform_2D(100, 100), beta, "averaged_sgd",
                                             {"mu": 0.5, "averaging": "ewma", "decay": 0.8, "precondition": 1.0}, target_update_every=0)
    kern = {8: "langevin2d_kernel<DoubleWell, Legendre2DStaticT<8,8>>", 32: "langevin2d_tc_kernel<32,32,6,1> (tensor-core contraction, folded derivative matrices)",
            64: "langevin2d_tc_kernel<64,64,4,1> (tensor-core contraction, folded derivative matrices)"}
    # the headline workload on the CUDA-core column sweep (the FP32-roofline kernel of round 1) and the 32 x 32 / 64 x 64 sweeps,
    # for the comparison with the tensor-core contraction that is the default now
    def headline_bias():
        return bias.VESBias_Ensemble(basis.LegendreBasis2D(15, 15, *MINMAX), target.Target_WellTempered_2D(100, 100, bias_factor=10.0), beta,
                                     "averaged_sgd", {"mu": args.mu, "averaging": "ewma", "decay": 0.8, "precondition": 1.0}, target_update_every=10)
    updater_leg("cfg2_cuda_cores", "configs[1] (the headline workload) with the step kernel forced onto the CUDA cores (pyves_set_tuning 6)",
                headline_bias, _lib.POT_DOUBLE_WELL_2D, (), MINMAX, 1000000, rank * 1000000, 1000000 * world, args.stride, 3,
                "langevin2d_kernel<DoubleWell, Legendre2DStaticT<16,16,const,4>> (CUDA cores, FFMA2 column sweep)", LEG_FLOPS[16], "weak", variant=6)
    for K, stride, vkern in ((32, 100, "langevin2d_kernel<DoubleWell, Legendre2DSmem<32,1,5>> (CUDA cores)"),
                             (64, 30, "langevin2d_kernel<DoubleWell, Legendre2DSmem<32,2,5>> (CUDA cores)")):
        updater_leg("cfg5_%dx%d_cuda_cores" % (K, K), "configs[4]: %dx%d Legendre on the CUDA cores (pyves_set_tuning 6), 10^6 walkers per GPU" % (K, K),
                    leg2d(K), _lib.POT_DOUBLE_WELL_2D, (), MINMAX, 1000000, rank * 1000000, 1000000 * world, stride, 2, vkern, LEG_FLOPS[K], "weak", variant=6)
    for K, stride in ((8, 500), (32, 100), (64, 30)):
        updater_leg("cfg5_%dx%d" % (K, K), "configs[4]: 2D double well + %dx%d Legendre, uniform target, 10^6 walkers per GPU" % (K, K), leg2d(K),
                    _lib.POT_DOUBLE_WELL_2D, (), MINMAX, 1000000, rank * 1000000, 1000000 * world, stride, 2, kern[K], LEG_FLOPS[K], "weak")
    updater_leg("cfg5_64x64_strong", "configs[4]: 2D double well + 64x64 Legendre, uniform target, 10^7 walkers sharded over the ranks", leg2d(64),
                _lib.POT_DOUBLE_WELL_2D, (), MINMAX, hi - lo, lo, 10 ** 7, 20, 2, kern[64], LEG_FLOPS[64], "strong")
    return out


def single_walker_rate(local, seed):
    """configs[0] as the reference runs it: ONE particle, static LegendreBasis1D(5) bias, 3 simulated dimensions, FP64, stepped
    through the C ABI -- steps/s of 2 x 10^5 steps in launches of 10^4 (the reference's Python loop reaches ~10^3 steps/s)."""
    import torch
    from pyves_b200 import _lib
    e = _lib.Ensemble(1, dims=3, seed=seed, precision="fp64", device=local)
    e.set_integrator(1.0, KB * TEMP, 20.0, 0.01)
    e.set_potential(_lib.POT_DOUBLE_WELL_2D)
    e.set_bias_legendre1d(0, 5, -2.5, 2.5, np.array([-11.8, 4.2, 13.1, -2.5, -9.0, 0.4]))
    e.set_state(np.array([[-1.4], [0.0], [0.0]]), np.zeros((3, 1)))
    e.step(1000)
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(20):
        e.step(10000)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t
    e.close()
    return {"value": 2e5 / dt, "unit": "steps/s", "note": "one walker, dims 3, FP64, general kernel, 20 launches of 10^4 steps (latency bound: "
            "one thread does the work); the reference's per-step Python/TorchForce loop is the torchforce_style figure of cpu_baseline"}


def tensor_roofline(basis, walker_steps_per_s):
    """Roofline keys of the step kernel whose contraction runs on the tensor cores (csrc/langevin_tc.cu).  `achieved` / `frac`
    stay the ALGORITHMIC flops of SURVEY 8(d) against the FP32 FMA peak -- the bound of the CUDA-core formulation the
    north-star names, which this kernel is allowed to beat because the 4 Kx Ky flops of the contraction left the FMA pipe;
    what bounds it instead is instruction issue (generation of the operand rows, x sweep, noise: profiles/r02_ncu_tc_step.md)
    and, for 64 x 64, shared-memory bandwidth.  `tensor` gives the executed tensor work against the FP16 dense peak."""
    kp = 16 if basis <= 16 else (32 if basis <= 32 else 64)
    executed = 12.0 * kp * kp                      # 3 MMAs (hi hi, hi lo, lo hi) x 2 N K flops per walker-step, N = 2 Kx (WA over WB), K = Ky
    peak16 = 1643.1
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            peak16 = float(json.load(fh)["bf16_tflops"])
    except (OSError, KeyError, ValueError):
        pass
    return {"bound": "instruction issue; the contraction runs on the tensor pipe (FP16 hi + lo split, tcgen05), so the FP32 FMA peak "
                     "is the bound of the CUDA-core formulation this kernel replaces, not of this kernel",
            "kernel": "langevin2d_tc_kernel",
            "tensor": {"executed_flops_per_walker_step": executed, "achieved": executed * walker_steps_per_s * 1e-12, "peak": peak16,
                       "unit": "TFLOP/s", "frac": executed * walker_steps_per_s * 1e-12 / peak16, "peak_source": "MEASURED_PEAKS.json bf16_tflops"}}


def workload_string(args):
    return ("configs[1]: 2D double well, %dx%d tensor Legendre bias on [-2.5,2.5]x[-2,2], well-tempered target (gamma 10, 100x100, "
            "refreshed every 10 updates), in-box ensemble averages, preconditioned averaged SGD (mu %g, EWMA 0.8) every %d steps, "
            "%d walkers per GPU" % (args.basis, args.basis, args.mu, args.stride, args.walkers))


def ncu_traffic(args):
    """DRAM traffic of the headline kernel cannot be read without a profiler: it comes from the committed ncu --set full
    capture named here (profiles/), scaled to this run's walker count, or is null."""
    path = os.path.join(ROOT, "profiles", "r02_traffic.json")
    try:
        with open(path) as fh:
            t = json.load(fh)
        if t.get("basis") == args.basis:
            return {"traffic": t["dram_bytes_per_walker"] * args.walkers,
                    "traffic_note": "NOT measured in this run: dram__bytes_read.sum + dram__bytes_write.sum per launch from the ncu --set full "
                                    "capture %s (%s), %g B per walker, scaled to %d walkers" % (t["source"], t["when"], t["dram_bytes_per_walker"], args.walkers)}
    except (OSError, KeyError, ValueError):
        pass
    return {"traffic": None, "traffic_note": "not measured in this run (needs ncu); algorithmic traffic is 32 B per walker per launch"}


def fes_rmse(bias, args):
    """FES read-out F = -V - kT ln p on the target grid against the analytic surface
    (x projection: ves/visualization.py:185-201; 2D: F = U), both restricted to F_ref <= 20 kT."""
    from pyves_b200 import fes
    from pyves_b200.potentials import DoubleWellPotential2D
    kT = KB_PY * TEMP
    pot = DoubleWellPotential2D()
    m = bias.target.mesh[0]
    F2 = bias.fes().reshape(bias.target.mesh) / kT
    cx = MINMAX[0] + (MINMAX[1] - MINMAX[0]) * (np.arange(m) + 0.5) / m
    cy = MINMAX[2] + (MINMAX[3] - MINMAX[2]) * (np.arange(bias.target.mesh[1]) + 0.5) / bias.target.mesh[1]
    Fx = -np.log(np.exp(-(F2 - F2.min())).sum(1))
    xs, F_ref = fes.projection_x(pot, TEMP, [-2.25, 2.25], [-2, 2])
    rm_x = fes.rmse_kt(np.interp(xs, cx, Fx - Fx.min()), F_ref, clip=20.0)
    gx, gy = np.meshgrid(cx, cy, indexing="ij")
    U = pot.potential(gx, gy) / kT
    U -= U.min()
    sel = U <= 20
    d = (F2 - F2[sel].min()) - U
    return {"rmse_x_kt": rm_x, "rmse_2d_kt": float(np.sqrt(np.mean(d[sel] ** 2))), "unit": "kT",
            "reference": "analytic F of DoubleWellPotential2D (x projection on a 200-point mesh; 2D surface), F_ref <= 20 kT"}


# ---------------------------------------------------------------------------------------------------
class CpuVes:
    """The headline workload on the host cores with the oracle's C port (gcc -O3, OpenMP) + numpy: the same VES iteration
    as the GPU arm -- `stride` Langevin steps of every walker, in-box first moments, preconditioned averaged-SGD update
    (mu, EWMA 0.8, scale (2i+1)(2j+1)), well-tempered target (gamma 10, 100 x 100 grid) refreshed every 10 updates."""

    def __init__(self, args, n):
        from oracle import cport, langevin, legendre, ves
        self.cport, self.ves, self.legendre = cport, ves, legendre
        cport.use_all_threads()
        self.args, self.n = args, n
        self.k = langevin.constants()
        K = args.basis
        self.spec = dict(kind="leg2d", deg_x=K - 1, deg_y=K - 1, minmax=MINMAX, weights=np.zeros((K, K)))
        self.opt = ves.AveragedSGD(np.zeros(K * K), mu=args.mu, decay=0.8)
        self.scale = np.outer(2 * np.arange(K) + 1, 2 * np.arange(K) + 1).ravel().astype(np.float64)
        self.beta = 1.0 / (KB_PY * TEMP)
        m = 100
        c = -1 + (2 * np.arange(m) + 1) / m
        gx, gy = np.meshgrid(c, c, indexing="ij")
        lv = np.polynomial.legendre.legvander
        self.basis_grid = (lv(gx.ravel(), K - 1)[:, :, None] * lv(gy.ravel(), K - 1)[:, None, :]).reshape(m * m, K * K)   # f_k at the target nodes
        self.p = np.full(m * m, 0.25)
        self.f_target = ves.target_average(self.basis_grid, self.p)
        self.n_updates = 0
        self.pos, self.vel = cpu_state(n, self.k, args.seed)
        self.step0 = 0

    def iteration(self):
        from oracle import potentials
        a = self.args
        m = self.cport.Model(potentials.DOUBLE_WELL_2D, (), self.spec, self.k, seed=a.seed, gid0=0)
        m.run(self.pos, self.vel, self.step0, a.stride)
        self.step0 += a.stride
        sw, sf = self.cport.moments_par(m, self.pos, inside_only=True)
        abar = self.opt.step((self.f_target - sf / sw) * self.scale)
        self.spec["weights"] = abar.reshape(self.spec["weights"].shape)
        self.n_updates += 1
        if self.n_updates % 10 == 0:
            q = self.ves.well_tempered_target(self.basis_grid @ abar, self.p, self.beta, 10.0)
            self.p = q / (q.sum() * 4.0 / q.size)
            self.f_target = self.ves.target_average(self.basis_grid, self.p)


def cpu_state(n, k, seed):
    from oracle import langevin, potentials
    en = lambda x, y: potentials.energy_grad_xy(potentials.DOUBLE_WELL_2D, x, y)[0]
    pos, ok = langevin.mc_placement(seed, 0, n, en, 1.0 / (KB_PY * TEMP), MINMAX)
    vel = langevin.maxwell_boltzmann(seed, 0, n, k, 2)
    return np.ascontiguousarray(pos.T), np.ascontiguousarray(vel.T)


def cpu_rate_probe(args):
    """walker-steps/s of the port on a 4096-walker probe (sizes the bounded samples)."""
    probe = CpuVes(args, 4096)
    stride, args.stride = args.stride, 50
    t = time.perf_counter()
    probe.iteration()
    dt = time.perf_counter() - t
    args.stride = stride
    return 4096 * 50 / dt


def torchforce_style_rate(args, steps=300):
    """What the reference executes per MD step for ONE particle: a TorchScript bias module
    evaluated forward + backward on the CPU, as openmm-torch's TorchForce does (ves/bias.py:52-58), plus the
    OpenMM-form Langevin update in numpy; one thread.  The module is this package's TorchScript export (the file a
    reference run would load from model_loc); OpenMM's own per-step overhead is not included (OpenMM is absent)."""
    import torch
    from pyves_b200.basis import LegendreBasis1D, LegendreBasis2D
    old_threads = torch.get_num_threads()
    torch.set_num_threads(1)
    out = {}
    kT, g, dt = KB * TEMP, 20.0, 0.01
    a = np.exp(-g * dt)
    b, c = (1 - a) / g, np.sqrt(kT * (1 - a * a))
    rng = np.random.default_rng(0)
    mods = {"legendre1d_d5": LegendreBasis1D(5, -2.5, 2.5, 'x', weights=rng.standard_normal(6)),
            "legendre2d_%dx%d" % (args.basis, args.basis): LegendreBasis2D(args.basis - 1, args.basis - 1, *MINMAX,
                                                                            weights=rng.standard_normal((args.basis, args.basis)) * 0.1)}
    for name, mod in mods.items():
        ts = mod.export_torchscript()
        x = np.array([[-1.4, 0.1, 0.0]])
        v = np.zeros((1, 3))
        for it in range(steps + 20):
            if it == 20:
                t0 = time.perf_counter()
            pos = torch.tensor(x, dtype=torch.float64, requires_grad=True)
            E = ts(pos).sum()
            E.backward()
            F = -pos.grad.numpy()
            F[0, 0] -= 10 * (4 * x[0, 0] ** 3 - 8 * x[0, 0] + 0.7)
            F[0, 1] -= 10 * x[0, 1]
            v = a * v + b * F + c * rng.standard_normal((1, 3))
            x = x + dt * v
        out[name] = steps / (time.perf_counter() - t0)
    # best-effort vectorised CPU (figure 2): the same TorchScript module on a batch of 10^5
    # particles, all cores, autograd for the forces, vectorised numpy integrator
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    nb, nst = 100000, 6
    ts = mods["legendre2d_%dx%d" % (args.basis, args.basis)].export_torchscript()
    x = np.concatenate([rng.uniform(-1.6, -1.2, (nb, 1)), rng.uniform(-0.3, 0.3, (nb, 1)), np.zeros((nb, 1))], 1)
    v = np.zeros((nb, 3))
    for it in range(nst + 2):
        if it == 2:
            t0 = time.perf_counter()
        pos = torch.tensor(x, dtype=torch.float64, requires_grad=True)
        ts(pos).sum().backward()
        F = -pos.grad.numpy()
        F[:, 0] -= 10 * (4 * x[:, 0] ** 3 - 8 * x[:, 0] + 0.7)
        F[:, 1] -= 10 * x[:, 1]
        v = a * v + b * F + c * rng.standard_normal((nb, 3))
        x = x + dt * v
    out["legendre2d_%dx%d_batched_1e5_all_cores" % (args.basis, args.basis)] = nb * nst / (time.perf_counter() - t0)
    torch.set_num_threads(old_threads)
    return out


def cpu_baseline(args, seconds=12.0):
    rate = cpu_rate_probe(args)
    n = int(max(2048, min(args.walkers, rate * seconds / args.stride)))
    cpu = CpuVes(args, n)
    t = time.perf_counter()
    cpu.iteration()
    dt = time.perf_counter() - t
    return {"value": n * args.stride / dt, "unit": "walker-steps/s", "cores": cpu.cport.num_threads(), "kind": "port",
            "sample": "%d of the %d walkers x %d steps + in-box moments + preconditioned averaged-SGD update, FP64, gcc -O3 OpenMP C port of the "
                      "same path (oracle/c/pyves_oracle.c); the reference itself (Python + OpenMM + TorchForce) cannot run here: "
                      "OpenMM absent" % (n, args.walkers, args.stride),
            "torchforce_style": dict(torchforce_style_rate(args), unit="walker-steps/s",
                                     note="TorchScript bias forward + backward per step on the CPU + numpy Langevin update: the "
                                          "per-step work of the reference's TorchForce path without OpenMM's own overhead; one "
                                          "particle on one thread (the reference's regime) and a 10^5-particle batch on all cores")}


def run_reference(args):
    """--impl reference: the CPU implementation of the headline workload (see CpuVes) on all host threads.  The FULL
    ensemble is run when the whole --steps/--warmup run fits --ref-budget seconds (same_config true), else a bounded sample
    of the walkers (every other element of the workload unchanged)."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    rate = cpu_rate_probe(args)
    iters = max(1, args.steps + args.warmup)
    n_fit = int(rate * args.ref_budget * 0.9 / (iters * args.stride))
    n = int(max(2048, min(args.walkers, n_fit)))
    full = n == args.walkers
    cpu = CpuVes(args, n)
    for _ in range(args.warmup):
        cpu.iteration()
    t = time.perf_counter()
    for _ in range(args.steps):
        cpu.iteration()
    dt = time.perf_counter() - t
    value = n * args.stride * args.steps / dt
    sample = ("%s: %d walkers x %d MD steps per bench step, FP64, %d host threads"
              % ("the full ensemble" if full else "bounded sample of the %d-walker workload" % args.walkers, n, args.stride, cpu.cport.num_threads()))
    out = {"impl": "reference", "metric": "walker-steps/s (2D double well, Legendre bias)", "value": value,
           "unit": "walker-steps/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
           "dtype": "f64", "data": "synthetic",
           "config": {"workload": workload_string(args), "sample": sample, "same_config": bool(full), "walkers_run": n},
           "cpu_baseline": {"value": value, "unit": "walker-steps/s", "cores": cpu.cport.num_threads(), "kind": "port",
                            "sample": sample + "; OpenMP C port of the path (oracle/c/pyves_oracle.c) + numpy target refresh -- the "
                                               "reference's own OpenMM + TorchForce stack is not installable here (no network)"},
           "e2e": {"value": value, "unit": "walker-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
