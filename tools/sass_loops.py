#!/usr/bin/env python
"""Summarise the instruction mix of the loops of each kernel in a cuobjdump -sass dump.

usage: cuobjdump -sass lib.so | python tools/sass_loops.py [kernel-name-substring]
For every kernel: each backward branch defines a loop [target, branch]; prints the opcode histogram
of the largest loops (the per-step body), which is what the roofline discussion in DESIGN.md uses.
"""
import collections
import re
import sys

pat = re.compile(r"/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)(.*?);")


def main():
    want = sys.argv[1] if len(sys.argv) > 1 else ""
    name = None
    ins = []
    kernels = []
    for line in sys.stdin:
        if "Function :" in line:
            if name is not None:
                kernels.append((name, ins))
            name = line.split("Function :")[1].strip()
            ins = []
            continue
        m = pat.search(line)
        if m and name is not None:
            ins.append((int(m.group(1), 16), m.group(2), m.group(3)))
    if name is not None:
        kernels.append((name, ins))
    for name, ins in kernels:
        if want not in name:
            continue
        print("==", name[:150], "instructions:", len(ins))
        loops = []
        for addr, op, rest in ins:
            if op.startswith("BRA"):
                m = re.search(r"0x([0-9a-f]+)", rest)
                if m and int(m.group(1), 16) < addr:
                    loops.append((addr - int(m.group(1), 16), int(m.group(1), 16), addr))
        loops.sort(reverse=True)
        for span, lo, hi in loops[:3]:
            hist = collections.Counter()
            for addr, op, rest in ins:
                if lo <= addr <= hi:
                    key = op.split(".")[0]
                    if key == "FFMA" and "UR" in rest:
                        key = "FFMA(ur)"
                    hist[key] += 1
            tot = sum(hist.values())
            print("  loop 0x%x..0x%x  %d instr:" % (lo, hi, tot),
                  " ".join("%s=%d" % kv for kv in hist.most_common(18)))


if __name__ == "__main__":
    main()
