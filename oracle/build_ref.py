"""TEST / BENCH INFRASTRUCTURE ONLY -- compile the reference's own environment modules into oracle/_ref/.

The reference is pure Python: its hot path is `MultiAgentEnv.step` of
environments/gym_flock_v2.py, gym_flock_uw.py, gym_flock_uw_discrete.py. This recipe byte-compiles
those three files FROM WHERE THEY LIE under /root/reference into `oracle/_ref/*.pyc.bin` (outputs only;
no source is copied, `oracle/_ref/` is git-ignored but travels to the GPU box with the snapshot, like
our own built `.so`). `oracle/ref_shim.py` loads the bytecode with importlib's sourceless loader when
/root/reference itself is absent, which is what lets `bench.py` time the UNMODIFIED reference step on
the GPU box's host cores (`cpu_baseline.reference_pytorch`, `--impl reference`).

    python -m oracle.build_ref         # also run by __graft_entry__.build() when /root/reference exists
"""
from __future__ import annotations

import os
import py_compile
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, "_ref")
REFERENCE_ROOT = os.environ.get("FLOCK_REFERENCE_ROOT", "/root/reference")
ENV_FILES = ("gym_flock_v2.py", "gym_flock_uw.py", "gym_flock_uw_discrete.py")


def build_ref(force: bool = False) -> bool:
    """Returns True when oracle/_ref/ holds the three compiled modules afterwards."""
    src_dir = os.path.join(REFERENCE_ROOT, "environments")
    if not all(os.path.isfile(os.path.join(src_dir, f)) for f in ENV_FILES):
        return all(os.path.isfile(os.path.join(REF_DIR, f + "c.bin")) for f in ENV_FILES)
    os.makedirs(REF_DIR, exist_ok=True)
    for f in ENV_FILES:
        src, out = os.path.join(src_dir, f), os.path.join(REF_DIR, f + "c.bin")
        if force or not os.path.isfile(out) or os.path.getmtime(out) < os.path.getmtime(src):
            # unchecked hash-based pyc: valid wherever the same interpreter version runs, whatever the mtime
            py_compile.compile(src, cfile=out, dfile=f"<reference>/environments/{f}", doraise=True,
                               invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
    with open(os.path.join(REF_DIR, "PYTHON_VERSION"), "w") as fh:
        fh.write("%d.%d\n" % sys.version_info[:2])
    return True


if __name__ == "__main__":
    print("oracle/_ref ready" if build_ref(force="--force" in sys.argv) else "reference tree not found; nothing built")
