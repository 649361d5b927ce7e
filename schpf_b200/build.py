"""Builds schpf_b200/_C/libschpf_b200.so from schpf_b200/csrc/*.cu with nvcc for sm_100a.

In-tree on purpose: the .so is git-ignored but travels with the repo snapshot
to the GPU box.  Run as ``python -m schpf_b200.build [--force]``.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
SRC_DIR = os.path.join(HERE, "csrc")
# experiments: SCHPF_BUILD_TAG=<name> SCHPF_NVCC_FLAGS="-DSWEEP_INTERLEAVE=0 ..." builds
# _C_<name>/libschpf_b200.so, which SCHPF_B200_LIB can then point the loader at
_TAG = os.environ.get("SCHPF_BUILD_TAG", "")
OUT_DIR = os.path.join(HERE, "_C" + ("_" + _TAG if _TAG else ""))
LIB = os.path.join(OUT_DIR, "libschpf_b200.so")
SOURCES = ["engine.cu", "sweep.cu", "sweep_lanes.cu", "sweep_f32.cu", "layout.cu", "ingest.cu", "dense.cu", "shims.cu"]
HEADERS = [os.path.join(SRC_DIR, "common.cuh"),
           os.path.join(HERE, "..", "include", "schpf_b200.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"] + os.environ.get("SCHPF_NVCC_FLAGS", "").split()


def _newer(a, b):
    return os.path.getmtime(a) > os.path.getmtime(b)


def build_library(force=False, verbose=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    objs, jobs = [], []
    for s in SOURCES:
        src = os.path.join(SRC_DIR, s)
        obj = os.path.join(OUT_DIR, s[:-3] + ".o")
        objs.append(obj)
        stale = force or not os.path.exists(obj) or _newer(src, obj) or any(_newer(h, obj) for h in HEADERS)
        if stale:
            cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        return cmd, r.returncode, r.stdout

    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1)) as ex:
            for cmd, rc, out in ex.map(run, jobs):
                if verbose or rc != 0:
                    sys.stderr.write(" ".join(cmd) + "\n" + out + "\n")
                if rc != 0:
                    raise RuntimeError("nvcc failed for %s" % cmd[-3])
    if jobs or not os.path.exists(LIB):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
