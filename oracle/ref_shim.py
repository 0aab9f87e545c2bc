"""TEST INFRASTRUCTURE ONLY -- loader for the UNMODIFIED reference environments on CPU.

This module is never imported by the product package. It exists so that
`tests/golden/make_golden.py` (and the optional live-reference tests, which skip
when `/root/reference` is absent) can execute the reference's own PyTorch code
and record golden input/output vectors that pin `oracle/flock_oracle.c`.

The reference modules (`environments/gym_flock_v2.py`, `gym_flock_uw.py`,
`gym_flock_uw_discrete.py`) import `gym` and `matplotlib` at module top and call
`.cuda()` on every allocation (e.g. gym_flock_v2.py:54-69). Neither package nor a
GPU exists in the build container, so the shim

  * registers minimal stand-ins for `gym`, `gym.spaces`, `matplotlib.pyplot`,
  * makes `Tensor.cuda()` / `Module.cuda()` the identity,
  * loads the file straight from `/root/reference/environments/` with importlib
    (no source is copied), from a scratch working directory because
    gym_flock_uw.py:13-14 / gym_flock_uw_discrete.py:13-14 create `./experiments`,
  * offers `NoiseInjector`, which replaces `torch.normal` (the actuation noise of
    gym_flock_uw_discrete.py:333-334) by `mean + <noise we supply>` so the same
    noise can be fed to the oracle / the CUDA kernels.

It cannot travel to the GPU box: nothing in `-m gpu` tests, `smoke()` or
`bench.py` imports it.
"""
from __future__ import annotations

import contextlib
import importlib.util
import os
import sys
import tempfile
import types

REFERENCE_ROOT = os.environ.get("FLOCK_REFERENCE_ROOT", "/root/reference")

_VARIANT_FILES = {
    "v2": "gym_flock_v2.py",
    "uw": "gym_flock_uw.py",
    "uwd": "gym_flock_uw_discrete.py",
}


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "environments", _VARIANT_FILES["v2"]))


def _install_stubs() -> None:
    import torch

    if "gym" not in sys.modules:
        gym = types.ModuleType("gym")
        spaces = types.ModuleType("gym.spaces")

        class Env:  # gym.Env stand-in: the reference only subclasses it
            pass

        class Box:
            def __init__(self, low, high, shape=None, dtype=None):
                self.low, self.high, self.shape = low, high, tuple(shape)

        class Discrete:
            def __init__(self, n):
                self.n = n
                self.shape = ()

        gym.Env = Env
        spaces.Box = Box
        spaces.Discrete = Discrete
        gym.spaces = spaces
        sys.modules["gym"] = gym
        sys.modules["gym.spaces"] = spaces
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self


_loaded: dict = {}


def load_reference(variant: str):
    """Return the reference module for `variant` in {"v2","uw","uwd"} (cached)."""
    if variant in _loaded:
        return _loaded[variant]
    if not reference_available():
        raise FileNotFoundError(f"reference tree not found under {REFERENCE_ROOT}")
    _install_stubs()
    path = os.path.join(REFERENCE_ROOT, "environments", _VARIANT_FILES[variant])
    spec = importlib.util.spec_from_file_location(f"_flock_reference_{variant}", path)
    mod = importlib.util.module_from_spec(spec)
    cwd = os.getcwd()
    scratch = tempfile.mkdtemp(prefix="flock_ref_")
    os.chdir(scratch)
    try:
        spec.loader.exec_module(mod)
    finally:
        os.chdir(cwd)
    _loaded[variant] = mod
    return mod


class NoiseInjector(contextlib.AbstractContextManager):
    """Replace `torch.normal(mean=..., std=...)` by `mean + noise[i]`.

    gym_flock_uw_discrete.py:333-334 draws the linear sample first, then the
    angular one; `noise` is a sequence of float32 tensors consumed in call order.
    """

    def __init__(self, noise):
        self.noise = list(noise)
        self.calls = 0

    def __enter__(self):
        import torch

        self._orig = torch.normal

        def fake_normal(mean=None, std=None, *a, **k):
            n = self.noise[self.calls]
            self.calls += 1
            return mean + n

        torch.normal = fake_normal
        return self

    def __exit__(self, *exc):
        import torch

        torch.normal = self._orig
        return False
