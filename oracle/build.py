"""Build the C oracle (oracle/hpf_oracle.c -> oracle/_build/libhpf_oracle.so).

TEST INFRASTRUCTURE ONLY -- see the header of hpf_oracle.c.  The reference is
pure Python/numba (no C sources), so there is no oracle/_ref to compile; the
real reference is used in the build container only, through
tests/golden/make_golden.py, to pin this oracle.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "hpf_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
OUT = os.path.join(OUT_DIR, "libhpf_oracle.so")


def build(force=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    if (not force and os.path.exists(OUT)
            and os.path.getmtime(OUT) >= os.path.getmtime(SRC)):
        return OUT
    # no -ffast-math: the oracle keeps IEEE evaluation order
    cmd = ["gcc", "-O2", "-fopenmp", "-fPIC", "-shared", "-std=c11",
           "-o", OUT, SRC, "-lm"]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
