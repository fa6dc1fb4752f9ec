#!/usr/bin/env python
"""Enqueue / execution cost of DeviceUpdater.update (device-resident coefficient update)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from pyves_b200.bias import DeviceUpdater
args = bench.parse_args()
beta = 1.0 / (bench.KB_PY * bench.TEMP)
from pyves_b200.potentials import _Potential
_Potential.quiet = True
bias = bench.make_bias(args, beta)
ens = bench.make_ensemble(args, 0, 0)
upd = DeviceUpdater(bias, ens)
def tick(): torch.cuda.synchronize(); return time.perf_counter()
T = {}
def add(k, v): T[k] = T.get(k, 0.0) + v
N = 40
for it in range(N + 5):
    if it == 5: T = {}
    t0 = tick(); ens.step(args.stride); e0 = time.perf_counter()
    t1 = tick(); m = ens.moments(); e1 = time.perf_counter()
    t2 = tick(); upd.update(m[0], m[1]); e2 = time.perf_counter()
    t3 = tick()
    add("step exec", t1 - t0); add("step enqueue", e0 - t0)
    add("moments exec", t2 - t1); add("moments enqueue", e1 - t1)
    add("update exec", t3 - t2); add("update enqueue", e2 - t2)
for k, v in T.items(): print("%-18s %.3f ms" % (k, v / N * 1e3))
# free-running loop without syncs
torch.cuda.synchronize(); t = time.perf_counter()
for it in range(N):
    ens.step(args.stride); m = ens.moments(); upd.update(m[0], m[1])
e = time.perf_counter() - t
torch.cuda.synchronize(); tot = time.perf_counter() - t
print("free-running: enqueue %.3f ms / iteration, total %.3f ms / iteration" % (e / N * 1e3, tot / N * 1e3))

def loop(tag, flush=None, events=False, monitor=False):
    mon = None
    if monitor:
        mon = bench.ClockMonitor(0); mon.start()
    torch.cuda.synchronize(); t = time.perf_counter()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    evs = []
    for it in range(N):
        if flush is not None: flush.zero_()
        if events:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True); a.record()
        ens.step(args.stride)
        if events: b.record(); evs.append((a, b))
        m = ens.moments(); upd.update(m[0], m[1])
    t1.record()
    torch.cuda.synchronize(); tot = time.perf_counter() - t
    if mon: mon.result()
    print("%-28s wall %.3f ms / iteration, events %.3f" % (tag, tot / N * 1e3, t0.elapsed_time(t1) / N))

fl = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
loop("plain")
loop("flush", flush=fl)
loop("flush+events", flush=fl, events=True)
loop("flush+events+monitor", flush=fl, events=True, monitor=True)
loop("monitor only", monitor=True)
