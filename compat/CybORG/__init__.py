"""Import alias: code written against the reference package layout (`from CybORG import CybORG`, `from CybORG.Agents import ...`,
`from CybORG.Agents.Wrappers import ...`, `from CybORG.Simulator.Scenarios import EnterpriseScenarioGenerator`) resolves to the
B200 engine's mirror of that API (cage_challenge_4_b200.wrappers) when THIS directory's parent (`compat/`) is on sys.path ahead
of a reference checkout.  With it the reference's own Evaluation/evaluation.py runs unchanged (tests/test_wrappers_gpu.py).
Only the names the CC4 hot path and its callers use exist here; see cage_challenge_4_b200/wrappers.py for what each one is.
"""
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
if _root not in _sys.path:
    _sys.path.insert(0, _root)

from cage_challenge_4_b200.wrappers import CybORG  # noqa: E402,F401

CYBORG_VERSION = "4.0 (cage_challenge_4_b200)"
