"""Build the in-tree native libraries (no JIT cache: the .so files travel with the repo snapshot).

  libecp_b200/lib/libecp_b200.so   host C layer (gcc) + sm_100a kernels (nvcc) - the product
  libecp_b200/lib/libecp.a         same objects as a static archive under the reference's archive name
  tests/hostcheck/libhostcheck.so  g++ build of csrc/ecp_math.h for CPU-side unit tests (test-only)
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
SO = os.path.join(LIBDIR, "libecp_b200.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC,-fopenmp", "-Xptxas", "-v"]
if os.environ.get("LIBECP_B200_NOFMAD"):  # A/B builds of the round-1 flag
    NVCC_FLAGS.insert(4, "-fmad=false")
CC_FLAGS = ["-O2", "-fPIC", "-Wall", "-Wno-comment", "-ffp-contract=off", "-std=gnu11", "-fopenmp"]


def _nvcc() -> str:
    for c in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def _run(cmd, log=None):
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if log is not None:
        log.append(p.stdout)
    if p.returncode != 0:
        sys.stderr.write(p.stdout)
        raise RuntimeError("command failed: " + " ".join(cmd))
    return p.stdout


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build_product(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    objdir = os.path.join(LIBDIR, "obj")
    os.makedirs(objdir, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    headers += [os.path.join(ROOT, "include", f) for f in os.listdir(os.path.join(ROOT, "include"))]
    objs = []
    log = []
    for name in ("tables.c", "builder.c", "api.c", "loaders.c"):
        src = os.path.join(CSRC, name)
        obj = os.path.join(objdir, name[:-2] + ".o")
        if force or _stale(obj, [src] + headers):
            _run(["gcc"] + CC_FLAGS + ["-c", src, "-o", obj], log)
        objs.append(obj)
    cu = os.path.join(CSRC, "ecp_cuda.cu")
    cuobj = os.path.join(objdir, "ecp_cuda.o")
    if force or _stale(cuobj, [cu] + headers):
        out = _run([_nvcc()] + NVCC_FLAGS + ["-c", cu, "-o", cuobj], log)
        with open(os.path.join(LIBDIR, "ptxas.log"), "w") as f:
            f.write(out)
    objs.append(cuobj)
    if force or _stale(SO, objs):
        _run([_nvcc(), "-shared", "-o", SO] + objs + ["-lm", "-lgomp"], log)
        _run(["ar", "rcs", os.path.join(LIBDIR, "libecp.a")] + objs, log)
    if verbose:
        print("".join(log))
    return SO


def build_hostcheck(force: bool = False) -> str:
    d = os.path.join(ROOT, "tests", "hostcheck")
    src = os.path.join(d, "hostcheck.cpp")
    so = os.path.join(d, "libhostcheck.so")
    if force or _stale(so, [src, os.path.join(CSRC, "ecp_math.h"), os.path.join(CSRC, "ecp_dev.h")]):
        _run(["g++", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-std=c++17", "-I", CSRC, src, "-o", so, "-lm"])
    return so


def build_oracle() -> None:
    """Compile oracle/'s C restatement and, when /root/reference is present, oracle/_ref (test infrastructure)."""
    _run(["make", "-C", os.path.join(ROOT, "oracle"), "all"])


if __name__ == "__main__":
    print(build_product(force="--force" in sys.argv, verbose=True))
