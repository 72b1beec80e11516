"""Times the UNMODIFIED Python reference (baseline/_ref/CybORG, placed there by baseline/install_ref.py) on the host cores.

One process per core (the reference is single-threaded Python), one CybORG environment each, driven through the reference's
own public API and stock code path exactly as a user would: EnterpriseScenarioGenerator(steps=500, FiniteStateRedAgent,
EnterpriseGreenAgent) -> CybORG(seed=int) -> BlueFlatWrapper.reset()/step(actions) with uniform-random action indices
(the workload of bench.py's own arm: full tick incl. the five observation vectors and the BlueRewardMachine reward; an
episode that ends is reset inside the timed region).  Nothing of this repository's engine is on that path; the only foreign
code are the import stand-ins in oracle/refshim for packages the image lacks (gym, gymnasium, pygame, prettytable, matplotlib,
ray, pettingzoo), which the reference touches at import time only.
"""
from __future__ import annotations

import multiprocessing as mp
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "baseline", "_ref")
SHIMS = os.path.join(ROOT, "oracle", "refshim")
N_ACTIONS = (82, 82, 82, 82, 242)

_env = None
_rng = None
_tick = 0


def available() -> bool:
    return os.path.isdir(os.path.join(REF_DIR, "CybORG"))


def _init(seed_base: int, policy: str):
    global _env, _rng, _tick, _policy
    for p in (REF_DIR, SHIMS):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    import numpy as np
    from CybORG import CybORG
    from CybORG.Agents import SleepAgent, EnterpriseGreenAgent, FiniteStateRedAgent
    from CybORG.Agents.Wrappers import BlueFlatWrapper
    from CybORG.Simulator.Scenarios import EnterpriseScenarioGenerator
    seed = seed_base
    sg = EnterpriseScenarioGenerator(blue_agent_class=SleepAgent, green_agent_class=EnterpriseGreenAgent,
                                     red_agent_class=FiniteStateRedAgent, steps=500)
    _env = BlueFlatWrapper(CybORG(sg, seed=seed))
    _env.reset()
    _rng = np.random.Generator(np.random.PCG64(seed))
    _tick, _policy = 0, policy


def _run(seconds: float):
    """Step this process's environment for `seconds`; returns (ticks, elapsed, resets)."""
    global _tick
    n = resets = 0
    t0 = time.perf_counter()
    while True:
        if _policy == "sleep":
            acts = {f"blue_agent_{i}": a for i, a in enumerate((49, 49, 49, 49, 145))}
        else:
            acts = {f"blue_agent_{i}": int(_rng.integers(0, k)) for i, k in enumerate(N_ACTIONS)}
        _, _, term, trunc, _ = _env.step(acts)
        n += 1
        _tick += 1
        if _tick >= 499 or all(term.values()) or all(trunc.values()):
            _env.reset()
            _tick = 0
            resets += 1
        dt = time.perf_counter() - t0
        if dt >= seconds:
            return n, dt, resets


def _worker(conn, seed_base: int, index: int, policy: str):
    _init(seed_base + index, policy)
    conn.send("ready")
    while True:
        msg = conn.recv()
        if msg is None:
            return
        conn.send(_run(float(msg)))


class ReferencePool:
    """`procs` worker processes, each holding one reference environment; every sample runs on all of them concurrently."""

    def __init__(self, procs: int | None = None, policy: str = "random", seed_base: int = 0):
        if not available():
            raise RuntimeError("baseline/_ref/CybORG is missing: run `python baseline/install_ref.py` in the build container")
        self.procs = procs or os.cpu_count() or 1
        ctx = mp.get_context("fork")
        self.workers = []
        for i in range(self.procs):
            a, b = ctx.Pipe()
            pr = ctx.Process(target=_worker, args=(b, seed_base, i, policy), daemon=True)
            pr.start()
            self.workers.append((pr, a))
        for _, a in self.workers:
            assert a.recv() == "ready"      # every worker has imported the reference and reset its environment

    def sample(self, seconds: float):
        """All workers step concurrently for `seconds`; returns (aggregate env ticks per second, ticks, wall seconds, resets)."""
        t0 = time.perf_counter()
        for _, a in self.workers:
            a.send(seconds)
        res = [a.recv() for _, a in self.workers]
        wall = time.perf_counter() - t0
        return sum(r[0] / r[1] for r in res), sum(r[0] for r in res), wall, sum(r[2] for r in res)

    def close(self):
        for pr, a in self.workers:
            a.send(None)
        for pr, _ in self.workers:
            pr.join(timeout=10)
