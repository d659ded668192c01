"""Build the CUDA library in-tree: `python -m sparrowrecsys_b200.build`.

nvcc cross-compiles for sm_100a without a GPU; the resulting
`sparrowrecsys_b200/libsrs_ctr.so` is git-ignored but travels with gpurun snapshots.
"""
from __future__ import annotations

import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libsrs_ctr.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.h")) + glob.glob(os.path.join(CSRC, "*.cuh")) \
        + glob.glob(os.path.join(os.path.dirname(HERE), "include", "*.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    nvcc = nvcc_path()
    common = [nvcc, *ARCH, "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
              "--expt-relaxed-constexpr", "--extended-lambda"]
    if verbose:
        common += ["-Xptxas", "-v"]
    procs = []
    objs = []
    for src in sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if (not force and os.path.exists(obj) and os.path.getmtime(obj) > max(
                [os.path.getmtime(src)] + [os.path.getmtime(h) for h in
                                           glob.glob(os.path.join(CSRC, "*.h")) +
                                           glob.glob(os.path.join(CSRC, "*.cuh")) +
                                           glob.glob(os.path.join(os.path.dirname(HERE), "include", "*.h"))])):
            continue
        procs.append((src, subprocess.Popen(common + ["-c", src, "-o", obj],
                                            stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    failed = False
    for src, p in procs:
        out = p.communicate()[0].decode()
        if p.returncode != 0:
            failed = True
            sys.stderr.write("nvcc failed on %s:\n%s\n" % (src, out))
        elif verbose or out.strip():
            sys.stderr.write(out)
    if failed:
        raise RuntimeError("CUDA build failed")
    tmp = LIB + ".tmp%d" % os.getpid()                   # link aside, then rename: a reader (or a gpurun
    subprocess.check_call([nvcc, *ARCH, "-shared", "-o", tmp, *objs])   # snapshot) never sees a half-written library
    os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
