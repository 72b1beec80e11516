"""The C-ABI library loads and exports every symbol include/cc4_b200.h declares (no compute without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "cc4_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cc4_[a-z_]+)\s*\(", text)))


def test_header_declares_the_documented_entry_points():
    syms = declared_symbols()
    for s in ("cc4_create", "cc4_reset", "cc4_step", "cc4_step_host", "cc4_rollout", "cc4_export_canon", "cc4_destroy"):
        assert s in syms


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    lib = ctypes.CDLL(os.path.join(ROOT, "cage_challenge_4_b200", "libcc4_b200.so"))
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} declared in include/cc4_b200.h but not exported"


def test_python_binding_lists_the_same_symbols():
    from cage_challenge_4_b200 import _native
    assert sorted(_native.EXPORTS) == declared_symbols()


def test_engine_refuses_to_run_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from cage_challenge_4_b200 import BatchedCC4, _native
    with pytest.raises(_native.NativeError):
        BatchedCC4(4)


def test_product_package_never_imports_the_oracle():
    """Neither the engine package nor the compat alias package may import, include or load anything of oracle/ (or baseline/)."""
    sources = []
    for pkg in (os.path.join(ROOT, "cage_challenge_4_b200"), os.path.join(ROOT, "compat")):
        for dirpath, _, files in os.walk(pkg):
            sources += [os.path.join(dirpath, f) for f in files if f.endswith((".py", ".cu", ".cuh", ".h"))]
    assert len(sources) > 20
    for path in sources:
        for line in open(path):
            code = line.strip()
            if code.startswith(("#include", "import ", "from ")) or "dlopen" in code or "CDLL" in code:
                assert "oracle" not in code and "ref_harness" not in code and "baseline" not in code, f"{path}: {code}"


def test_host_expansion_of_bit_rows_matches_numpy():
    """cc4_expand_obs_bits (host-only, what cc4_step_host's thread pool runs): bit rows -> compact and padded byte rows."""
    import numpy as np
    from cage_challenge_4_b200 import _native
    L = _native.load()
    g = np.random.default_rng(0)
    n = 257
    rows = g.integers(0, 2**32, (n, 20), dtype=np.uint64).astype(np.uint32)
    rows[:, 19] = g.integers(0, 3, n)
    bits = np.unpackbits(rows[:, :19].copy().view(np.uint8), axis=1, bitorder="little")[:, :578]
    off = [0, 92, 184, 276, 368, 578]
    want = bits.copy()
    want[:, off[:5]] = rows[:, 19:20].astype(np.uint8)
    got = np.full((n, 578), 7, dtype=np.uint8)
    assert L.cc4_expand_obs_bits(rows.ctypes.data, got.ctypes.data, n, 1) == 0
    assert np.array_equal(got, want)
    padded = np.full((n, 5, 210), 7, dtype=np.uint8)
    assert L.cc4_expand_obs_bits(rows.ctypes.data, padded.ctypes.data, n, 0) == 0
    for b in range(5):
        ln = off[b + 1] - off[b]
        assert np.array_equal(padded[:, b, :ln], want[:, off[b]:off[b + 1]]) and not padded[:, b, ln:].any()


def test_unpack_bits_matches_the_host_expansion():
    """engine.unpack_bits (torch, any device) and cc4_expand_obs_bits (C, host) are the same map."""
    import numpy as np
    import torch
    from cage_challenge_4_b200 import _native
    from cage_challenge_4_b200.engine import unpack_bits
    L = _native.load()
    g = np.random.default_rng(1)
    n = 33
    rows = g.integers(0, 2**32, (n, 20), dtype=np.uint64).astype(np.uint32)
    rows[:, 19] = g.integers(0, 3, n)
    want = np.zeros((n, 578), dtype=np.uint8)
    assert L.cc4_expand_obs_bits(rows.ctypes.data, want.ctypes.data, n, 1) == 0
    got = unpack_bits(torch.from_numpy(rows.view(np.int32)))
    assert np.array_equal(got.numpy(), want)


def test_compat_package_resolves_the_reference_import_paths():
    """compat/CybORG: `from CybORG import CybORG`, `CybORG.Agents...`, `CybORG.Simulator.Scenarios...` (the imports of the
    reference's Evaluation/evaluation.py:5-7 and example_submission/submission.py:1-5) resolve to the mirrored API."""
    import subprocess
    import sys
    code = ("from CybORG import CybORG, CYBORG_VERSION\n"
            "from CybORG.Agents import SleepAgent, EnterpriseGreenAgent, FiniteStateRedAgent, BaseAgent, RandomSelectRedAgent, DiscoveryFSRed, cc4BlueRandomAgent\n"
            "from CybORG.Agents.Wrappers import BlueFixedActionWrapper, BlueFlatWrapper, BlueEnterpriseWrapper, EnterpriseMAE\n"
            "from CybORG.Agents.Wrappers.EnterpriseMAE import EnterpriseMAE as E2\n"
            "from CybORG.Simulator.Scenarios import EnterpriseScenarioGenerator\n"
            "from CybORG.Shared.Enums import TernaryEnum\n"
            "import cage_challenge_4_b200.wrappers as w\n"
            "assert CybORG is w.CybORG and E2 is w.EnterpriseMAE and EnterpriseScenarioGenerator is w.EnterpriseScenarioGenerator\n"
            "sg = EnterpriseScenarioGenerator(blue_agent_class=SleepAgent, green_agent_class=EnterpriseGreenAgent, red_agent_class=FiniteStateRedAgent, steps=500)\n"
            "print('ok', CYBORG_VERSION)\n")
    env = dict(os.environ, PYTHONPATH=os.path.join(ROOT, "compat"))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd="/tmp")
    assert out.returncode == 0 and out.stdout.startswith("ok"), out.stderr[-800:]


def test_production_kernels_carry_no_tape_code():
    """The device code is compiled into two namespaces (csrc/cc4_engine.cu): `cc4` with the tape paths (parity replay) and the
    reset kernels, `cc4p` = k_step / k_rollout of production mode without them.  The shipped library must hold both sets of
    kernels, and no tape-mode function may exist in the production namespace."""
    import shutil
    import subprocess
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    from cage_challenge_4_b200 import _native
    out = subprocess.run(["cuobjdump", "-symbols", _native.LIB_PATH], capture_output=True, text=True).stdout
    names = [l.split()[-1] for l in out.splitlines() if "STT_FUNC" in l]
    for rv in range(3):
        assert any(f"N4cc4p6k_stepILi{rv}E" in n for n in names), f"production k_step<{rv}> missing"
        assert any(f"N3cc46k_stepILi{rv}E" in n for n in names), f"tape-mode k_step<{rv}> missing"
        assert any(f"N4cc4p9k_rolloutILi{rv}E" in n for n in names), f"production k_rollout<{rv}> missing"
    assert any("N3cc412k_reset_fast" in n for n in names) and any("N3cc47k_reset" in n for n in names)
    tape_in_prod = [n for n in names if "N4cc4p" in n and "tape" in n]
    assert not tape_in_prod, tape_in_prod
    assert any("N3cc4" in n and "tape_next" in n for n in names), "the tape reader must exist in the parity namespace"
