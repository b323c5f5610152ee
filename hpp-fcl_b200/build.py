"""In-tree build of the CUDA extension (sm_100a only) with nvcc.

The product is a plain C-ABI shared library (include/hppfcl_b200.h): no torch
types cross the boundary, so it is built with nvcc directly rather than through
torch.utils.cpp_extension.
"""
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
_SO = os.path.join(_HERE, "libhppfcl_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    # bit-exact parity with the FP64 reference semantics: no FMA contraction
    # (the oracle is compiled -ffp-contract=off); fp64 div/sqrt are IEEE already.
    "-fmad=false",
    "-Xcompiler", "-fPIC", "-shared",
    "-Xptxas", "-v",
]


def library_path():
    return _SO


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def sources():
    return sorted(os.path.join(_CSRC, f) for f in os.listdir(_CSRC) if f.endswith(".cu"))


def _deps():
    out = []
    for root, _, files in os.walk(_CSRC):
        out += [os.path.join(root, f) for f in files if f.endswith((".cu", ".cuh", ".h"))]
    out.append(os.path.join(_HERE, "..", "include", "hppfcl_b200.h"))
    return out


def build_extension(force=False, verbose=False):
    """Compile csrc/*.cu -> libhppfcl_b200.so (cross-compiles without a GPU)."""
    if not force and os.path.exists(_SO):
        t = os.path.getmtime(_SO)
        if all(os.path.getmtime(d) <= t for d in _deps()):
            return _SO
    cmd = [_nvcc()] + NVCC_FLAGS + ["-I", os.path.join(_HERE, "..", "include"), "-I", _CSRC,
                                     "-o", _SO] + sources()
    env = dict(os.environ)
    # the image exports CXX=/opt/gcc/bin/g++ (a wrapper); nvcc wants the system host compiler
    cmd += ["-ccbin", "/usr/bin/g++"] if os.path.exists("/usr/bin/g++") else []
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)
    log = os.path.join(_HERE, "build.log")
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + res.stdout)
    if verbose or res.returncode != 0:
        print(res.stdout)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed (see %s)" % log)
    return _SO
