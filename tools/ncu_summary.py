#!/usr/bin/env python
"""Summarise ncu output into markdown for profiles/.

  python tools/ncu_summary.py raw  <ncu --page raw --csv file>      -> per-kernel table of the key metrics
  python tools/ncu_summary.py list <ncu --csv launch list (gpu__time_duration.sum)> [skip_launches]
                                                                      -> per-kernel time share
"""
import collections
import csv
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum", "sm__cycles_elapsed.avg",
]
STALL = "smsp__average_warps_issue_stalled_"


def raw(path):
    rows = [r for r in csv.reader(open(path, newline="")) if len(r) > 20]
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    for r in data:
        print("\n## `%s`\n" % r[ix["Kernel Name"]][:170])
        print("| metric | value | unit |\n|---|---|---|")
        for k in KEYS:
            if k in ix and r[ix[k]] != "":
                print("| %s | %s | %s |" % (k, r[ix[k]], units[ix[k]]))
        st = sorted(((float(r[i] or 0), h[len(STALL):].replace("_per_issue_active.ratio", "")) for h, i in ix.items()
                     if h.startswith(STALL) and h.endswith("_per_issue_active.ratio")), reverse=True)
        print("\nStall reasons (average warps per issue-active cycle): " + ", ".join("%s %.2f" % (n, v) for v, n in st[:7]))


def launches(path, skip=0):
    rows = [r for r in csv.reader(open(path, newline="")) if len(r) > 10]
    hdr = rows[0]
    k, v = hdr.index("Kernel Name"), hdr.index("Metric Value")
    data = rows[1:][skip:]
    tot = collections.OrderedDict()
    cnt = collections.Counter()
    for r in data:
        name = r[k].split("(")[0][:90]
        tot[name] = tot.get(name, 0.0) + float(r[v].replace(",", "")) / 1e3
        cnt[name] += 1
    s = sum(tot.values())
    print("| kernel | launches | total us | share |\n|---|---|---|---|")
    for n, t in sorted(tot.items(), key=lambda kv: -kv[1]):
        print("| `%s` | %d | %.1f | %.3f |" % (n, cnt[n], t, t / s))
    print("| total | %d | %.1f | 1 |" % (sum(cnt.values()), s))


if __name__ == "__main__":
    if sys.argv[1] == "raw":
        raw(sys.argv[2])
    else:
        launches(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 0)
