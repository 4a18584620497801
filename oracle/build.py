"""Compile the C oracle (test infrastructure) into oracle/_build/liboracle.so.

MDAnalysis (where the reference's arithmetic lives) has no sources under
/root/reference, so there is no `oracle/_ref` to build: the reference path is
"unbuildable here" and the oracle is a restatement ("port").
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "rdf_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
OUT = os.path.join(OUT_DIR, "liboracle.so")

# strict IEEE on purpose: no -ffast-math, no FMA contraction
CFLAGS = ["-O2", "-fopenmp", "-ffp-contract=off", "-fno-fast-math",
          "-shared", "-fPIC", "-Wall", "-Wextra", "-std=c11"]


def build(force=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    if (not force and os.path.exists(OUT)
            and os.path.getmtime(OUT) >= os.path.getmtime(SRC)):
        return OUT
    cmd = ["gcc"] + CFLAGS + [SRC, "-o", OUT, "-lm"]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
