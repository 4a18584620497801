"""Build librdfk.so (sm_100a only) in-tree with nvcc.

    python -m pmda_b200.csrc.build [--force] [--verbose]

The .so stays next to the sources (git-ignored, but it travels to the GPU box
with the gpurun snapshot).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
INCLUDE = os.path.join(ROOT, "include")
SOURCES = [os.path.join(HERE, "rdfk.cu")]
OUT = os.path.join(HERE, "librdfk.so")
#: the same sources with -DRDFK_CHECK (device-side guard rails on the pair
#: kernel's queue / log bounds); one GPU test runs the parity cases through it
OUT_CHECK = os.path.join(HERE, "librdfk_check.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "--fmad=false",          # every FMA in the kernels is written explicitly
    "-Xcompiler", "-fPIC", "-Xcompiler", "-pthread", "-shared",
    "-I", INCLUDE,
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.sep not in cand or os.path.exists(cand)):
            return cand
    return "nvcc"


def needs_build(out=None):
    out = out or OUT
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    deps = SOURCES + [os.path.join(HERE, "hbond.cuh"),
                      os.path.join(HERE, "hosthash.inl"),
                      os.path.join(INCLUDE, "rdfk.h"), os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, out=None, defines=()):
    """`out` / `defines`: experimental variants next to the default library
    (selected at run time with RDFK_LIB)"""
    if out is None and not force and not needs_build():
        return OUT
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) \
        + ["-D" + d for d in defines] + SOURCES + ["-o", out or OUT]
    subprocess.check_call(cmd)
    return out or OUT


def build_check(force=False):
    """the RDFK_CHECK variant next to the release library"""
    if not force and not needs_build(OUT_CHECK):
        return OUT_CHECK
    return build(out=OUT_CHECK, defines=("RDFK_CHECK",))


if __name__ == "__main__":
    _out = [a[6:] for a in sys.argv if a.startswith("--out=")]
    _def = [a[2:] for a in sys.argv if a.startswith("-D")]
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv,
                out=_out[0] if _out else None, defines=_def))
