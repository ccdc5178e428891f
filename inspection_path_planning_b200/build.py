"""Builds libipp.so in-tree with nvcc for sm_100a (no JIT cache, no torch extension machinery).

    python -m inspection_path_planning_b200.build [--force] [--verbose]

Translation units whose fp64 results must equal the reference arithmetic bit for bit (exact.cu,
plan_reduce.cu, evaluator.cu) are compiled with -fmad=false; host code everywhere with -ffp-contract=off.
"""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libipp.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-ffp-contract=off,-Wall,-Wno-unused-function",
          "-I", os.path.join(HERE, "..", "include")]
# (source, extra flags)
UNITS = [
    ("exact.cu", ["-fmad=false"]),
    ("plan_reduce.cu", ["-fmad=false"]),
    ("evaluator.cu", ["-fmad=false"]),
    ("lbvh.cu", []),
    ("visibility.cu", []),
    ("visibility_bins.cu", []),
    ("pose_dedupe.cu", []),
    ("comm.cu", []),
    ("ipp_api.cu", []),
    ("mesh_io.cpp", []),
    ("planner.cpp", []),
]


def _nvcc():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found")
    return nvcc


def _host_cxx():
    # the image's CXX wrapper (/opt/gcc/bin/g++) does not link libstdc++ into shared objects; use the plain one
    for c in ("/usr/bin/g++", shutil.which("g++")):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("g++ not found")


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    nvcc, cxx = _nvcc(), _host_cxx()
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    headers.append(os.path.join(HERE, "..", "include", "ipp.h"))
    headers.append(os.path.abspath(__file__))
    jobs = []
    objs = []
    for src, extra in UNITS:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, os.path.splitext(src)[0] + ".o")
        objs.append(o)
        if force or _newer(o, [s] + headers):
            cmd = [nvcc, "-ccbin", cxx] + ARCH + COMMON + extra + ["-c", s, "-o", o]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        return cmd, r

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        for cmd, r in ex.map(run, jobs):
            if verbose or r.returncode:
                sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
            if r.returncode:
                raise RuntimeError("nvcc failed for " + cmd[-3])
    if jobs or force or _newer(LIB, objs):
        cmd = [nvcc, "-ccbin", cxx, "-shared", "-o", LIB] + objs + ["-lcudart_static", "-ldl", "-lrt", "-lpthread"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode:
            sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode:
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
