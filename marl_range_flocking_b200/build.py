"""Build libflock_b200.so (sm_100a) in-tree with nvcc. `python -m marl_range_flocking_b200.build`.

The shared library lands next to this file so that it travels to the GPU box with the repo
snapshot. nvcc cross-compiles for sm_100a without a GPU.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ_DIR = os.path.join(HERE, "_build")
LIB_PATH = os.path.join(HERE, "libflock_b200.so")
SOURCES = ["flock_small_v2p.cu", "flock_small_v2e.cu", "flock_small_v2pn.cu", "flock_small_v2en.cu", "flock_small_uw.cu", "flock_small_uwn.cu", "flock_small_uwd.cu",
           "flock_small_uwdn.cu", "flock_small.cu", "flock_tiled.cu", "flock_actor.cu", "flock_rnn_actor.cu", "flock_qnet.cu", "flock_gru_tc.cu", "flock_api.cu"]
FAST_MATH_OK = {"flock_actor.cu", "flock_rnn_actor.cu", "flock_qnet.cu", "flock_gru_tc.cu"}
HEADERS = ["flock_device.cuh", "flock_small_impl.cuh", "flock_launch.h", "flock_tc.cuh", os.path.join("..", "..", "include", "flock_b200.h")]

# -fmad=false: no implicit FMA contraction (canonical arithmetic); explicit fmaf()/fma()
# calls still emit FFMA/DFMA. Precise division / sqrt, denormals kept (the nvcc defaults, stated).
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
              "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.isfile(nvcc):
        raise RuntimeError("nvcc not found; cannot build libflock_b200.so")
    return nvcc


def _stale(target: str, deps) -> bool:
    if not os.path.isfile(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ_DIR, exist_ok=True)
    nvcc = _nvcc()
    hdrs = [os.path.normpath(os.path.join(CSRC, h)) for h in HEADERS] + [os.path.abspath(__file__)]
    jobs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ_DIR, src[:-3] + ".o")
        if force or _stale(o, [s] + hdrs):
            flags = NVCC_FLAGS
            if src in FAST_MATH_OK:      # tensor-core policy kernel: not part of the bit-exact env path
                flags = [f for f in NVCC_FLAGS if f != "-fmad=false"]
            cmd = [nvcc, *flags, "-c", s, "-o", o]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            jobs.append(cmd)
    with cf.ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as ex:
        for res in ex.map(lambda c: subprocess.run(c, capture_output=True, text=True), jobs):
            if verbose or res.returncode != 0:
                sys.stderr.write(res.stdout + res.stderr)
            if res.returncode != 0:
                raise RuntimeError("nvcc failed: " + " ".join(res.args))
    objs = [os.path.join(OBJ_DIR, s[:-3] + ".o") for s in SOURCES]
    if force or jobs or _stale(LIB_PATH, objs):
        cmd = [nvcc, "-shared", "-o", LIB_PATH, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
               "-cudart", "static", "-Xcompiler", "-fPIC"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError("link failed")
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
