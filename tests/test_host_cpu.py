"""CPU: host-side logic, the C ABI surface, and the multi-rank plumbing (gloo, world_size 2).

No compute call is made here: without a GPU the library must refuse loudly instead of falling back.
"""
import ctypes
import os
import re
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="session")
def lib():
    from marl_range_flocking_b200 import build as flock_build
    flock_build.build()
    from marl_range_flocking_b200 import load_library
    return load_library()


def _header():
    with open(os.path.join(ROOT, "include", "flock_b200.h")) as f:
        return f.read()


def test_library_exports_every_declared_symbol(lib):
    names = re.findall(r"FLOCK_API[^;(]*?\b(flock_\w+)\s*\(", _header())
    assert len(names) >= 17
    out = subprocess.run(["nm", "-D", "--defined-only", lib._name], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (flock_\w+)", out))
    assert set(names) <= exported, sorted(set(names) - exported)
    for n in names:
        assert getattr(lib, n) is not None


def test_only_the_c_abi_is_exported(lib):
    out = subprocess.run(["nm", "-D", "--defined-only", lib._name], capture_output=True, text=True, check=True).stdout
    syms = re.findall(r" [TW] (\S+)", out)
    assert all(s.startswith("flock_") for s in syms), [s for s in syms if not s.startswith("flock_")][:5]


def test_library_contains_sm100a_code(lib):
    out = subprocess.run(["cuobjdump", "--list-elf", lib._name], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_struct_layouts_match_header():
    from marl_range_flocking_b200._lib import BUFFER_FIELDS, FlockBuffers, FlockCfg
    h = _header()
    cfg_body = re.search(r"typedef struct flock_cfg_t \{(.*?)\} flock_cfg_t;", h, re.S).group(1)
    cfg_body = re.sub(r"/\*.*?\*/", "", cfg_body, flags=re.S)
    fields = re.findall(r"\b(?:int32_t|float|uint64_t)\s+(\w+);", cfg_body)
    assert fields == [f[0] for f in FlockCfg._fields_]
    assert ctypes.sizeof(FlockCfg) == 8 * 4 + 10 * 4 + 8
    buf_body = re.search(r"typedef struct flock_buffers_t \{(.*?)\} flock_buffers_t;", h, re.S).group(1)
    buf_body = re.sub(r"/\*.*?\*/", "", buf_body, flags=re.S)
    names = []
    for decl in re.findall(r"\b(?:float|int32_t|uint8_t|uint32_t|int64_t|uint64_t)\s+([^;]+);", buf_body):
        names += [n.strip().lstrip("*") for n in decl.split(",")]
    assert names == list(BUFFER_FIELDS)
    assert ctypes.sizeof(FlockBuffers) == 8 * len(BUFFER_FIELDS)


def test_config_validation_errors_come_back_through_the_abi(lib):
    from marl_range_flocking_b200._lib import FlockCfg
    h = ctypes.c_void_p()

    def cfg(**kw):
        base = dict(variant=0, num_envs=4, num_agents=10, k=4, rigid_boundary=0, periodic=1, obs_hist=1, env_offset=0,
                    boundary=50.0, range_lo=0.0, reset_hi=50.0, heading_hi=4.7, sensor_range=14.0,
                    collision_distance=2.5, reset_collision_distance=2.5, max_linear_velocity=2.5, act_noise_std=0.0,
                    range_noise_std=0.0, seed=1)
        base.update(kw)
        return FlockCfg(**base)

    for bad, msg in ((dict(num_agents=4), "k+1"), (dict(k=0), "k must be"), (dict(k=9, num_agents=20), "k must be"),
                     (dict(variant=3), "variant"), (dict(num_envs=0), "num_envs"), (dict(obs_hist=4), "obs_hist"),
                     (dict(variant=1, periodic=1, obs_hist=4), "periodic"), (dict(boundary=0.0), "boundary"),
                     (dict(num_agents=9000), "FLOCK_MAX_AGENTS")):
        rc = lib.flock_create(ctypes.byref(cfg(**bad)), 0, ctypes.byref(h))
        assert rc == -1 and msg in lib.flock_last_error().decode(), (bad, lib.flock_last_error())
        assert not h.value


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU refusal")
def test_no_gpu_means_loud_failure_not_fallback(lib):
    from marl_range_flocking_b200 import VecEnv
    from marl_range_flocking_b200._lib import FlockCfg
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        VecEnv("v2", 4, 10, 4, 2.5)
    h = ctypes.c_void_p()
    c = FlockCfg(0, 4, 10, 4, 0, 1, 1, 0, 50.0, 0.0, 50.0, 4.7, 14.0, 2.5, 2.5, 2.5, 0.0, 0.0, 1)
    assert lib.flock_create(ctypes.byref(c), 0, ctypes.byref(h)) == -4          # FLOCK_E_NO_DEVICE
    assert b"no CPU fallback" in lib.flock_last_error()
    assert lib.flock_step(None, None, 0.1, None, None) == -1                   # null handle, no crash


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "marl_range_flocking_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
                assert "libflock_oracle" not in src and "orc_" not in src, f
                assert not re.search(r'#include\s+"[^"]*oracle', src), f


def test_spaces_match_reference_surface():
    from marl_range_flocking_b200.spaces import Box, Discrete
    b = Box(low=-1.5, high=1.5, shape=(2,))
    assert b.shape == (2,) and b.shape[0] == 2 and b.contains(b.sample())
    assert float(b.low[0]) == -1.5 and float(b.high[1]) == 1.5
    d = Discrete(4)
    assert d.n == 4 and d.contains(d.sample()) and not d.contains(4)
    assert Box(low=0, high=50, shape=(10, 4)).shape == (10, 4)


def test_shard_range_partitions_the_envs():
    from marl_range_flocking_b200.dist import shard_range
    for total in (1, 7, 64, 4096, 4099):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == total
            for (o0, c0), (o1, _) in zip(spans, spans[1:]):
                assert o0 + c0 == o1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def test_stats_dict_decodes_fixed_point():
    from marl_range_flocking_b200.dist import stats_dict
    fx = int(round(-12.5 * 10 * 2**32))      # two episodes, N = 10 agents, summed return -12.5 per agent
    d = stats_dict(torch.tensor([2, 300, fx, 5, 1, 0, 0, 0], dtype=torch.int64), agents=10)
    assert d["episodes"] == 2 and d["mean_episode_length"] == 150.0
    assert abs(d["mean_episode_return"] - (-6.25)) < 1e-9 and d["reset_gave_up"] == 1


_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["FLOCK_ROOT"])
from marl_range_flocking_b200.dist import allreduce_stats, shard_range, stats_dict
dist.init_process_group("gloo")
r, w = dist.get_rank(), dist.get_world_size()
off, cnt = shard_range(101, w, r)
# each rank contributes statistics proportional to its shard; negative fixed-point returns included
local = torch.tensor([cnt, 10 * cnt, -(off + 1) * 2**32, r, 0, 0, 0, 0], dtype=torch.int64)
tot = allreduce_stats(local)
spans = [shard_range(101, w, i) for i in range(w)]
want = [101, 1010, -sum(o + 1 for o, _ in spans) * 2**32, sum(range(w)), 0, 0, 0, 0]
assert tot.tolist() == want, (tot.tolist(), want)
assert local[0].item() == cnt      # input untouched
d = stats_dict(tot, agents=1)
assert d["episodes"] == 101
dist.barrier()
dist.destroy_process_group()
print("rank", r, "ok")
"""


def test_stats_allreduce_two_ranks_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = dict(os.environ, FLOCK_ROOT=ROOT, CUDA_VISIBLE_DEVICES="")
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                         capture_output=True, text=True, env=env, timeout=240)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert res.stdout.count("ok") == 2


def test_bench_reference_arm_prints_one_json_line():
    import json
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "5",
                          "--warmup", "1"], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["value"] > 0 and j["cpu_baseline"]["kind"] in ("reference", "port")
    assert j["cpu_baseline"]["port"]["value"] > 0          # the C port is always timed next to the reference
    if j["cpu_baseline"]["kind"] == "reference":           # oracle/_ref present: the unmodified reference step
        assert "unmodified reference" in j["cpu_baseline"]["sample"] and j["value"] < j["cpu_baseline"]["port"]["value"]
    assert j["e2e"]["h2d_bytes_per_step"] == 0 and j["unit"] == "agent-steps/s"
    # the config object is the GPU arm's, key for key (the driver compares them)
    sys.path.insert(0, ROOT)
    import bench
    assert j["config"] == bench.config_of("cfg2", bench.WORKLOADS["cfg2"])


def test_compiled_reference_loads_without_the_source_tree():
    """oracle/build_ref.py byte-compiles the reference env modules into oracle/_ref; the shim must be able to
    load them when /root/reference is absent (what bench.py does on the GPU box)."""
    from oracle import build_ref
    if not build_ref.build_ref():
        pytest.skip("no reference tree and no compiled copy on this machine")
    code = ("import os, sys; sys.path.insert(0, %r); os.environ['FLOCK_REFERENCE_ROOT'] = '/nonexistent';"
            "from oracle import ref_shim; assert not ref_shim.reference_available() and ref_shim.compiled_reference_available();"
            "import torch; m = ref_shim.load_reference('uw'); e = m.MultiAgentEnv(agents=6, k=3, collision_distance=0.5, range_start=(0, 100));"
            "o = e.reset(); o, r, d, _ = e.step(torch.rand(6, 2)); assert o.shape == (6, 4, 3) and r.shape == (6, 1)") % ROOT
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300,
                         env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
    assert res.returncode == 0, res.stderr[-2000:]
