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

The reference TREE does not travel to the GPU box. What does travel is `oracle/_ref/`
(git-ignored build output of `oracle/build_ref.py`: the three environment modules byte-compiled
from where they lie under /root/reference). When /root/reference is absent the shim loads that
bytecode with importlib's sourceless loader, so the CPU-baseline legs of `bench.py` can time the
UNMODIFIED reference step on the box's host cores. `-m gpu` tests and `smoke()` never import it.
"""
from __future__ import annotations

import contextlib
import importlib.machinery
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


_REF_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def reference_available() -> bool:
    """The reference source tree is present (build container)."""
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "environments", _VARIANT_FILES["v2"]))


def compiled_reference_available() -> bool:
    """oracle/_ref/ holds the byte-compiled reference modules for THIS interpreter version."""
    try:
        with open(os.path.join(_REF_DIR, "PYTHON_VERSION")) as fh:
            if fh.read().strip() != "%d.%d" % sys.version_info[:2]:
                return False
    except OSError:
        return False
    return all(os.path.isfile(os.path.join(_REF_DIR, f + "c.bin")) for f in _VARIANT_FILES.values())


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
        # the learner networks finish their constructors with self.to(cuda device)
        # (learners/maddpg_shared_critic/ddpg_network.py:128-130): a no-op on a CPU-only machine
        if not getattr(torch.nn.Module.to, "_flock_shim", False):
            _orig_to = torch.nn.Module.to

            def _to(self, *a, **k):
                dev = a[0] if a else k.get("device")
                if isinstance(dev, (str, torch.device)) and torch.device(dev).type == "cuda":
                    return self
                return _orig_to(self, *a, **k)

            _to._flock_shim = True
            torch.nn.Module.to = _to


_loaded: dict = {}


def load_reference(variant: str):
    """Return the reference module for `variant` in {"v2","uw","uwd"} (cached)."""
    if variant in _loaded:
        return _loaded[variant]
    name = f"_flock_reference_{variant}"
    if reference_available():
        path = os.path.join(REFERENCE_ROOT, "environments", _VARIANT_FILES[variant])
        spec = importlib.util.spec_from_file_location(name, path)
    elif compiled_reference_available():
        path = os.path.join(_REF_DIR, _VARIANT_FILES[variant] + "c.bin")
        spec = importlib.util.spec_from_loader(name, importlib.machinery.SourcelessFileLoader(name, path))
    else:
        raise FileNotFoundError(f"reference tree not found under {REFERENCE_ROOT} and no compiled copy in {_REF_DIR}")
    _install_stubs()
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


def load_learner_module(relpath: str):
    """Load a learner module of the reference tree by path relative to its root, e.g.
    "learners/vdn/net.py" (test infrastructure: pins policies.py / replay.py to the reference's own classes).
    Only possible where the source tree exists (the build container)."""
    key = "learner:" + relpath
    if key in _loaded:
        return _loaded[key]
    path = os.path.join(REFERENCE_ROOT, relpath)
    if not os.path.isfile(path):
        raise FileNotFoundError(path)
    _install_stubs()
    name = "_flock_reference_" + relpath.replace("/", "_").replace(".py", "")
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    cwd = os.getcwd()
    scratch = tempfile.mkdtemp(prefix="flock_ref_")
    os.makedirs(os.path.join(scratch, "tmp"), exist_ok=True)   # Actor.__init__ does os.mkdir("tmp\\ddpg") style paths
    os.chdir(scratch)
    try:
        spec.loader.exec_module(mod)
    finally:
        os.chdir(cwd)
    mod._flock_scratch = scratch
    _loaded[key] = mod
    return mod


@contextlib.contextmanager
def scratch_cwd(mod=None):
    """Run reference constructors that create directories in the cwd from a scratch directory."""
    cwd = os.getcwd()
    d = getattr(mod, "_flock_scratch", None) or tempfile.mkdtemp(prefix="flock_ref_")
    os.chdir(d)
    try:
        yield d
    finally:
        os.chdir(cwd)


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
