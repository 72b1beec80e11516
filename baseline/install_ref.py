#!/usr/bin/env python
"""Install the UNMODIFIED reference into baseline/_ref/ so that `bench.py --impl reference` can time it on the GPU box.

baseline/_ref/ is git-ignored (reference sources never enter this repository's history) but not gpurun-ignored: it travels
to the GPU box with the snapshot, like the built .so files.  Runs in the build container only (needs /root/reference).

Step 1 is the prescribed command:
    python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target baseline/_ref <copy of /root/reference>
(from a copy under /tmp because the build writes into the source tree; --no-deps because gym / ray / ... are not in the
wheelhouse).  Outcome, recorded in DESIGN.md: it succeeds but installs ONLY `cyborg-4.0.dist-info`, because the reference's
setup.py declares `py_modules=[]` and no packages (setup.py:15-21) — the reference is meant to be used from its checkout
(`pip install -e .`).  Step 2 therefore places the package directory itself (CybORG/, byte-for-byte, minus the test suites under Tests/ and
__pycache__) next to the dist-info, which is what an editable install exposes on sys.path.  The import stand-ins for the
packages that are absent from this image (gym, gymnasium, pygame, prettytable, matplotlib, ray, pettingzoo: touched at
import time only) are the repo's own oracle/refshim/ and are put on sys.path by the caller.
"""
import filecmp
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("CC4_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")


def install(verbose: bool = True) -> bool:
    if not os.path.isdir(os.path.join(REF, "CybORG")):
        return False
    pkg = os.path.join(DST, "CybORG")
    if os.path.isdir(pkg) and filecmp.cmp(os.path.join(pkg, "env.py"), os.path.join(REF, "CybORG", "env.py"), shallow=False):
        return True
    os.makedirs(DST, exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp:
        src = os.path.join(tmp, "reference")
        shutil.copytree(REF, src, ignore=shutil.ignore_patterns("__pycache__", "documentation"))
        r = subprocess.run([sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--find-links", "/opt/wheelhouse",
                            "--upgrade", "--target", DST, src], capture_output=True, text=True)
        if verbose:
            print("pip install --target baseline/_ref:", "ok" if r.returncode == 0 else f"rc={r.returncode}", "->", sorted(os.listdir(DST)))
    if os.path.isdir(pkg):
        shutil.rmtree(pkg)
    shutil.copytree(os.path.join(REF, "CybORG"), pkg, ignore=shutil.ignore_patterns("__pycache__", "test_cc4", "test_python"))   # env.py:23 imports CybORG.Tests.utils
    if verbose:
        print("placed the package directory:", pkg)
    return True


if __name__ == "__main__":
    sys.exit(0 if install() else 1)
