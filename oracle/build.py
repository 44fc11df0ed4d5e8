"""Compile oracle/dvr_ref.c -> oracle/_build/liboracle_dvr.so with gcc (+OpenMP).

Test infrastructure: building the checker is not using it.  No fast-math and no FMA
contraction so the fp64 traversal follows the reference's operation order exactly.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "liboracle_dvr.so")
SRC = os.path.join(HERE, "dvr_ref.c")


def build(force=False):
    os.makedirs(OUT, exist_ok=True)
    if (not force and os.path.exists(LIB)
            and os.path.getmtime(LIB) >= os.path.getmtime(SRC)):
        return LIB
    cmd = ["gcc", "-O2", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off",
           "-std=c11", "-o", LIB, SRC, "-lm"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("gcc failed on oracle/dvr_ref.c:\n" + r.stdout)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
