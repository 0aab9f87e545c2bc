"""Per-CTA timeline of the step kernels (developer tool; evidence for DESIGN 5.1 / 5.2a).

Build the instrumented library once (CPU box is fine):   python tools/build_variant.py timeline -DFLOCK_TIMELINE --only=flock_
then, on a GPU:                                           python tools/cta_timeline.py [cfg2|cfg3|cfg4|uw2048|uwd2048|v22048]
The instrumented kernels record %globaltimer at CTA entry, after griddepcontrol.wait and at exit, the SM id and -- for
the pruned large-swarm kernel -- the slowest warp's cycles in the shared and in the one-at-a-time pass, the rows taken
out of the shared pass and the boxes opened. Prints the distribution for the LAST launch of a short run."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ.setdefault("FLOCK_LIBRARY_PATH", os.path.join(ROOT, "marl_range_flocking_b200", "_ab", "libflock_timeline.so"))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from marl_range_flocking_b200 import VecEnv, _lib  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
lib = _lib.load_library()
lib.flock_debug_timeline.restype = ctypes.c_int
lib.flock_debug_timeline.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
dev = torch.device("cuda", 0)


def dump(env, n):
    buf = (ctypes.c_ulonglong * (8 * n))()
    rc = lib.flock_debug_timeline(env._h, buf, n)
    assert rc == 0, rc
    return np.array(buf, dtype=np.uint64).reshape(n, 8).astype(np.int64)


def pct(v):
    return "p0 %.2f p50 %.2f p90 %.2f max %.2f" % tuple(np.percentile(v, [0, 50, 90, 100]))


if what in bench.WORKLOADS and not what == "cfg5":
    w = dict(bench.WORKLOADS[what])
    ring, _ = bench.ring_size(w)
    envs, acts = bench.build_ring(w, w["E"], ring, dev, env_offset=0)
    g, _ = bench.capture(envs, acts, 10 * ring, 0)
    g.replay()
    g.replay()
    torch.cuda.synchronize()
    last = envs[(10 * ring - 1) % ring]
    n = (last.num_envs + 2 * (32 // last.num_particles) - 1) // (2 * (32 // last.num_particles))
    a = dump(last, n)
    t0 = a[:, 0].min()
    print(f"{what}: {n} CTAs of the last step of a graph-replayed stream (us after the first CTA entry)")
    print("  CTA entry                 ", pct((a[:, 0] - t0) / 1e3))
    print("  griddepcontrol.wait done  ", pct((a[:, 1] - t0) / 1e3))
    print("  CTA exit                  ", pct((a[:, 2] - t0) / 1e3))
    print("  work phase per CTA (us)   ", pct((a[:, 2] - a[:, 1]) / 1e3))
    per_sm = np.bincount(a[:, 3], minlength=148)
    print("  CTAs per SM: min %d max %d" % (per_sm.min(), per_sm.max()))
    last_exit = np.array([((a[a[:, 3] == s, 2] - t0) / 1e3).max() if per_sm[s] else 0.0 for s in range(len(per_sm))])
    for c in sorted(set(per_sm[per_sm > 0])):
        sel = last_exit[per_sm == c]
        print("  SMs with %2d CTAs: %3d, last exit mean %.2f max %.2f us" % (c, len(sel), sel.mean(), sel.max()))
else:
    variant, k = {"uw2048": ("uw", 3), "uwd2048": ("uwd", 4), "v22048": ("v2", 8), "cfg5": ("v2", 8)}[what]
    env = VecEnv(variant, 64, 2048, k, 0.05, range_start=(0, 2000), sensor_range=100.0, seed=3, reset_collision_distance=0.05)
    env.reset()
    acts = [env.random_actions(i) for i in range(2)]
    n = 1024
    for i in range(32):
        env.step(acts[i & 1])
        a = dump(env, n)
        if i in (3, 15, 17, 22, 31):
            t0 = a[:, 0].min()
            st, en = (a[:, 0] - t0) / 1e3, (a[:, 2] - t0) / 1e3
            o = np.argsort(-(en - st))[:4]
            print(f"{what} step {i}: kernel span {en.max():.1f} us, CTAs entering later than 5 us: {int((st > 5).sum())}; CTA duration mean {(en - st).mean():.1f} "
                  f"max {(en - st).max():.1f} us; slowest warp, shared pass: mean {a[:, 4].mean():.0f} max {a[:, 4].max()} cycles; one-at-a-time pass: mean "
                  f"{a[:, 5].mean():.0f} max {a[:, 5].max()} cycles; rows taken out: {a[:, 6].sum()} (max {a[:, 6].max()} per CTA); slowest CTAs (env, tile, us, "
                  f"shared-pass cycles, rows out, boxes):", [(int(x // 16), int(x % 16), round(float(en[x] - st[x]), 1), int(a[x, 4]), int(a[x, 6]), int(a[x, 7])) for x in o])
