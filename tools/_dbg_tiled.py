import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from marl_range_flocking_b200 import VecEnv, _lib
lib = _lib.load_library()
variant, k = sys.argv[1], int(sys.argv[2])
env = VecEnv(variant, 64, 2048, k, 0.05, range_start=(0, 2000), sensor_range=100.0, seed=3, reset_collision_distance=0.05)
env.reset()
acts = [env.random_actions(i) for i in range(2)]
n = 1024
buf = (ctypes.c_ulonglong * (8 * n))()
for i in range(32):
    env.step(acts[i & 1])
    lib.flock_debug_tiled_dump(buf, n, 1)
    if i in (15, 16, 17, 22):
        a = np.array(buf, dtype=np.uint64).reshape(n, 8).astype(np.int64)
        g0, g1, main, farc, nfar, ev, sm, tot = [a[:, j] for j in range(8)]
        t0 = g0.min()
        st, en = (g0 - t0) / 1e3, (g1 - t0) / 1e3
        o = np.argsort(-(en - st))[:4]
        print(f"{variant} step {i}: span {en.max():.1f} us; late starters {int((st > 5).sum())}; CTA dur mean {(en-st).mean():.1f} max {(en-st).max():.1f}; main-pass cycles (slowest warp) mean {main.mean():.0f} max {main.max():.0f}; far-pass cycles mean {farc.mean():.0f} max {farc.max():.0f}; far rows total {nfar.sum()} max/CTA {nfar.max()}; slowest CTAs (dur us, main, far, nfar, boxes):",
              [(int(x // 16), int(x % 16), round(float(en[x]-st[x]),1), int(main[x]), int(nfar[x]), int(ev[x])) for x in o], 'CTAs with bad-hint rows', int(((nfar // 1000) % 1000 > 0).sum()), 'with loose thr', int((nfar // 1000000 > 0).sum()), flush=True)
