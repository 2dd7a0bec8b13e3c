"""
In-tree build of libpypde_b200.so for sm_100a (nvcc cross-compiles without a GPU).

    python -m pypde_b200.build [--force]

Sources: pypde_b200/csrc/*.cu -> pypde_b200/_lib/*.o -> pypde_b200/_lib/libpypde_b200.so
(the CUDA runtime is linked statically, so the library loads on a box without a
driver; every compute call then fails with PDE_ERR_CUDA instead of falling back).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "_lib")
LIB = os.path.join(OUT, "libpypde_b200.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-I" + os.path.join(ROOT, "include"), "-I" + CSRC]
COMMON += os.environ.get("PDE_NVCC_EXTRA", "").split()      # developer experiments (e.g. -DPDE_SW_KC=8)
# banded.cu keeps the Fortran operation order (bit parity with the oracle): no FMA contraction
PER_FILE = {"banded.cu": ["--fmad=false"], "batched.cu": ["--fmad=false"]}


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(OUT, exist_ok=True)
    nvcc = _nvcc()
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(ROOT, "include", "pypde_b200.h"))
    objs, cmds = [], []
    for src in sources():
        s = os.path.join(CSRC, src)
        o = os.path.join(OUT, src[:-3] + ".o")
        objs.append(o)
        if force or _stale(o, [s] + headers):
            cmds.append([nvcc] + ARCH + COMMON + PER_FILE.get(src, []) + ["-c", s, "-o", o])
    rebuilt = bool(cmds)
    if cmds:
        from concurrent.futures import ThreadPoolExecutor

        def run(cmd):
            if verbose:
                print(" ".join(cmd))
            subprocess.check_call(cmd)
        with ThreadPoolExecutor(max_workers=min(len(cmds), os.cpu_count() or 4)) as ex:
            list(ex.map(run, cmds))
    if rebuilt or force or _stale(LIB, objs):
        cmd = [nvcc] + ARCH + ["-shared", "-o", LIB] + objs
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
