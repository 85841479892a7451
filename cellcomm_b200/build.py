"""Build libcellcomm_b200.so in-tree with nvcc for sm_100a.

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.  There is no
fallback: if the library is missing and cannot be built, importing the ops fails loudly.
"""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "csrc", "build")
LIB = os.path.join(HERE, "libcellcomm_b200.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "--expt-relaxed-constexpr", "--extended-lambda",
    "-Xcompiler", "-fPIC",
    "-I", INCLUDE,
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libcellcomm_b200.so")


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith((".cu", ".cpp")))


def _deps_mtime():
    m = 0.0
    for f in os.listdir(CSRC):
        if f.endswith((".cuh", ".h")):
            m = max(m, os.path.getmtime(os.path.join(CSRC, f)))
    for f in os.listdir(INCLUDE):
        m = max(m, os.path.getmtime(os.path.join(INCLUDE, f)))
    return m


def is_stale():
    if not os.path.exists(LIB):
        return True
    lib_m = os.path.getmtime(LIB)
    if _deps_mtime() > lib_m:
        return True
    return any(os.path.getmtime(os.path.join(CSRC, s)) > lib_m for s in sources())


def build_library(force=False, verbose=False):
    """Compile every csrc/*.cu|*.cpp for sm_100a and link the shared library."""
    if not force and not is_stale():
        return LIB
    nvcc = _nvcc()
    os.makedirs(BUILD, exist_ok=True)
    hdr_m = _deps_mtime()

    def compile_one(src):
        obj = os.path.join(BUILD, os.path.splitext(src)[0] + ".o")
        src_path = os.path.join(CSRC, src)
        if (not force and os.path.exists(obj)
                and os.path.getmtime(obj) > max(os.path.getmtime(src_path), hdr_m)):
            return obj
        cmd = [nvcc] + NVCC_FLAGS + \
              ["-c", src_path, "-o", obj]
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{res.stdout}\n{res.stderr}")
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, sources()))
    tmp = LIB + ".tmp"
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", tmp] + objs + \
          ["-lpthread"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True))
