"""A/B builds: libipp.so with extra -D switches on the K4 translation units -> _ab/<name>.so (the tree's own build is untouched).

    python tools/build_variant.py <name> -DIPP_X=1 [-DIPP_Y=2 ...]
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from inspection_path_planning_b200 import build as b  # noqa: E402

name, defs = sys.argv[1], sys.argv[2:]
b.build()
out_dir = os.path.join(ROOT, "_ab", "obj_" + name)
os.makedirs(out_dir, exist_ok=True)
objs = []
for src, extra in b.UNITS:
    o = os.path.join(b.OBJ, os.path.splitext(src)[0] + ".o")
    if src in ("visibility_bins.cu", "visibility.cu"):
        o = os.path.join(out_dir, os.path.splitext(src)[0] + ".o")
        cmd = [b._nvcc(), "-ccbin", b._host_cxx()] + b.ARCH + b.COMMON + extra + defs + ["-c", os.path.join(b.CSRC, src), "-o", o]
        if "--verbose" in os.environ.get("IPP_BUILD", ""):
            cmd.insert(1, "-Xptxas=-v")
        subprocess.check_call(cmd)
    objs.append(o)
lib = os.path.join(ROOT, "_ab", name + ".so")
subprocess.check_call([b._nvcc(), "-ccbin", b._host_cxx(), "-shared", "-o", lib] + objs + ["-lcudart_static", "-ldl", "-lrt", "-lpthread"])
print(lib)
