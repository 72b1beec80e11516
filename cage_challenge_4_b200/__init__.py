"""cage_challenge_4_b200 — B200-native batched step engine for CAGE Challenge 4 (CybORG).

Only the hot path is here: the multi-agent CC4 `step()` (reference: CybORG/Simulator/SimulationController.py:211)
for N concurrent episodes as sm_100a CUDA kernels behind a C-ABI (`include/cc4_b200.h`), plus the host-side mirror of
the reference's wrapper API (`wrappers.py`).  There is no CPU fallback: importing `engine` without the built
extension raises.
"""
from .engine import BatchedCC4, N_BLUE, OBS_LEN, MASK_LEN, N_ACTIONS, SLEEP_INDEX  # noqa: F401

__all__ = ["BatchedCC4", "N_BLUE", "OBS_LEN", "MASK_LEN", "N_ACTIONS", "SLEEP_INDEX"]
