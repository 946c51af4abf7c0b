"""Builds libsiftcuda.so (hand-written CUDA for sm_100a) in-tree with nvcc.

    python -m siftmetal_b200.build [--force]

The .so is git-ignored but travels with the repo snapshot to the GPU box. No JIT cache, no
torch extension machinery: the product is a plain C-ABI shared library (include/siftcuda.h).
"""
import concurrent.futures
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libsiftcuda.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-std=c++17", "-O3", "-lineinfo", "-Xcompiler", "-fPIC", "-prec-div=true", "-prec-sqrt=true"]
# Bit-exact stages are compiled with contraction off: only explicit fmaf() fuses.
SOURCES = {
    "pyramid.cu": ["-fmad=false"],
    "detect.cu": ["-fmad=false"],
    "describe.cu": (["-DSIFT_DEBUG_DESC"] if os.environ.get("SIFT_DEBUG_DESC") else []),
    "match.cu": [],
    "geometry.cu": ["-Xcompiler", "-ffp-contract=off", "-fmad=false"],
    "capi.cu": [],
}
HEADERS = ["common.cuh", "dev_math.cuh", "scan.cuh", "capi_match.inc", os.path.join(ROOT, "include", "siftcuda.h")]


def nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found")
    return exe


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(BUILD, exist_ok=True)
    hdrs = [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in HEADERS]
    cc = nvcc()
    jobs = []
    objs = []
    for src, extra in SOURCES.items():
        obj = os.path.join(BUILD, src.replace(".cu", ".o"))
        objs.append(obj)
        path = os.path.join(CSRC, src)
        if force or _stale(obj, [path] + hdrs):
            cmd = [cc] + ARCH + COMMON + extra + ["-Xptxas", "-v", "-c", path, "-o", obj]
            jobs.append((src, cmd))

    def run(job):
        src, cmd = job
        r = subprocess.run(cmd, capture_output=True, text=True)
        return src, r

    logs = {}
    with concurrent.futures.ThreadPoolExecutor(max_workers=4) as ex:
        for src, r in ex.map(run, jobs):
            logs[src] = r.stderr
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
    if jobs or force or _stale(LIB, objs):
        cmd = [cc] + ARCH + ["-shared", "-o", LIB] + objs + ["-cudart", "static", "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    if verbose:
        for src, log in logs.items():
            print(f"==== {src}\n{log}")
        with open(os.path.join(BUILD, "ptxas.log"), "w") as f:
            for src, log in logs.items():
                f.write(f"==== {src}\n{log}\n")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
