#!/usr/bin/env python
"""Build libpyves_b200.so (sm_100a) in-tree:  python pyves_b200/csrc/build.py [--force]

Plain nvcc, no torch: the C ABI of include/pyves_b200.h is the product boundary.  Translation units
compile in parallel; the step-kernel instantiations are split by (precision, group).
"""
import concurrent.futures
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libpyves_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC"]

UNITS = [("api.cu", []), ("misc_kernels.cu", []), ("moments.cu", []), ("cov_tc.cu", []), ("comm.cu", []), ("langevin_tc.cu", [])]
for grp in range(4):
    for t in ("float", "double"):
        UNITS.append(("langevin_inst.cu", ["-DPV_T=%s" % t, "-DPV_GROUP=%d" % grp]))


def sources_digest():
    h = hashlib.sha256()
    for root in (HERE, os.path.join(HERE, "..", "..", "include")):
        for name in sorted(os.listdir(root)):
            if name.endswith((".cu", ".cuh", ".h", ".py")):
                with open(os.path.join(root, name), "rb") as fh:
                    h.update(name.encode())
                    h.update(fh.read())
    return h.hexdigest()


def compile_unit(unit):
    src, defs = unit
    tag = src.replace(".cu", "") + "".join(d.replace("-D", "_").replace("=", "-") for d in defs)
    obj = os.path.join(OBJ, tag + ".o")
    cmd = [NVCC] + FLAGS + defs + ["-c", os.path.join(HERE, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s %s:\n%s" % (src, defs, r.stderr))
    return obj


def build(force=False, verbose=True):
    os.makedirs(OBJ, exist_ok=True)
    stamp = os.path.join(OBJ, "digest")
    digest = sources_digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == digest:
        if verbose:
            print("libpyves_b200.so is up to date")
        return LIB
    with concurrent.futures.ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        objs = list(ex.map(compile_unit, UNITS))
    cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stderr)
    with open(stamp, "w") as fh:
        fh.write(digest)
    if verbose:
        print("built", LIB)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
