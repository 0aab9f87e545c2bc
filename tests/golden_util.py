"""Helpers shared by the oracle-vs-golden (CPU) and CUDA-vs-golden (GPU) tests."""
from __future__ import annotations

import glob
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def traj_files():
    return sorted(glob.glob(os.path.join(GOLDEN_DIR, "traj_*.npz")))


def load_traj(path):
    z = np.load(path, allow_pickle=False)
    d = {k: z[k] for k in z.files}
    d["variant"] = str(d["variant"])
    cfg = {}
    for k in list(d):
        if k.startswith("cfg_"):
            v = d.pop(k)
            cfg[k[4:]] = tuple(v.tolist()) if v.ndim else v.item()
    d["cfg"] = cfg
    return d


def load_edges():
    z = np.load(os.path.join(GOLDEN_DIR, "edge_cases.npz"), allow_pickle=False)
    cases = []
    for i in range(int(z["num_cases"])):
        pre = f"c{i}_"
        c, cfg = {}, {}
        for k in z.files:
            if not k.startswith(pre):
                continue
            name = k[len(pre):]
            v = z[k]
            if name.startswith("cfg_"):
                cfg[name[4:]] = tuple(v.tolist()) if v.ndim else v.item()
            elif name in ("label", "variant"):
                c[name] = str(v)
            else:
                c[name] = v
        c["cfg"] = cfg
        cases.append(c)
    return cases


def env_kwargs(variant, cfg):
    """Reference ctor kwargs -> kwargs common to OracleEnv and the product VecEnv."""
    kw = dict(agents=int(cfg["agents"]), k=int(cfg.get("k", 4)),
              collision_distance=float(cfg.get("collision_distance", 3)),
              range_start=tuple(cfg.get("range_start", (0, 100))),
              sensor_range=float(cfg.get("sensor_range", 7)),
              rigid_boundary=bool(cfg.get("rigid_boundary", False)))
    return kw


def close(a, b, rtol=1e-5, atol=1e-6):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.abs(a - b) <= atol + rtol * np.abs(b)


def torus_close(a, b, B, rtol=1e-5, atol=1e-5):
    """Positions compared modulo the wrap rule (B and 0.001 are the two images of the wall)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    d = np.abs(a - b)
    d = np.minimum(d, np.abs(B - d))
    return d <= atol + rtol * B + 0.0011
