"""In-tree build of the CUDA engine: nvcc -> qiskit-aer_b200/libb200sv.so (sm_100a only).

The .so is git-ignored but travels to the GPU box with the repo snapshot.  Translation units are compiled in
parallel into build/*.o (only the stale ones) and linked with nvcc.
"""
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libb200sv.so")
SOURCES = ["gates.cu", "tile.cu", "reduce.cu", "planner.cu", "sharded.cu", "api.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC"]
HEADERS = [os.path.join(CSRC, "common.cuh"), os.path.join(HERE, "..", "include", "b200sv.h")]


def _mtime(p):
    return os.path.getmtime(p) if os.path.exists(p) else 0.0


def _stale():
    t = _mtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + HEADERS
    return t == 0.0 or any(_mtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OBJ, exist_ok=True)
    hdr_t = max(_mtime(h) for h in HEADERS + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))])

    def compile_one(src):
        s, o = os.path.join(CSRC, src), os.path.join(OBJ, src[:-3] + ".o")
        if not force and _mtime(o) > max(_mtime(s), hdr_t):
            return o
        cmd = [nvcc] + NVCC_FLAGS + ["-c", s, "-o", o]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
        return o

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a"] + objs + ["-o", LIB]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    import sys
    build(force="--incremental" not in sys.argv, verbose=True)
