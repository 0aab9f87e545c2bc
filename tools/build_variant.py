"""Developer A/B: build libflock_b200 with extra -D defines into marl_range_flocking_b200/_ab/libflock_<tag>.so
(git-ignored; travels with gpurun). usage: python tools/build_variant.py <tag> [-DNAME=VAL ...] [--only=prefix]
(--only: recompile just the sources starting with `prefix`, default flock_small; the rest comes from the regular _build/)
Run a variant with FLOCK_LIBRARY_PATH=marl_range_flocking_b200/_ab/libflock_<tag>.so python bench.py ..."""
import concurrent.futures as cf
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from marl_range_flocking_b200 import build as B

tag = sys.argv[1]
defs = [a for a in sys.argv[2:] if not a.startswith("--only=")]
only = ([a[7:] for a in sys.argv[2:] if a.startswith("--only=")] or ["flock_small"])[0]
B.build()
obj_dir = os.path.join(B.HERE, "_build_" + tag)
out_dir = os.path.join(B.HERE, "_ab")
os.makedirs(obj_dir, exist_ok=True)
os.makedirs(out_dir, exist_ok=True)
jobs = []
mine = [s for s in B.SOURCES if s.startswith(only)]
for src in mine:
    flags = [f for f in B.NVCC_FLAGS if not (src in B.FAST_MATH_OK and f == "-fmad=false")]
    jobs.append([B._nvcc(), *flags, *defs, "-c", os.path.join(B.CSRC, src), "-o", os.path.join(obj_dir, src[:-3] + ".o")])
with cf.ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
    for res in ex.map(lambda c: subprocess.run(c, capture_output=True, text=True), jobs):
        if res.returncode != 0:
            sys.exit(res.stdout + res.stderr)
objs = [os.path.join(obj_dir if s in mine else B.OBJ_DIR, s[:-3] + ".o") for s in B.SOURCES]
lib = os.path.join(out_dir, f"libflock_{tag}.so")
res = subprocess.run([B._nvcc(), "-shared", "-o", lib, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static",
                      "-Xcompiler", "-fPIC"], capture_output=True, text=True)
if res.returncode != 0:
    sys.exit(res.stdout + res.stderr)
print(lib)
