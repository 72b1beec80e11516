"""Distributional fixtures from the live reference (TEST INFRASTRUCTURE; needs /root/reference, build container only).

Runs N full reference episodes (EnterpriseScenarioGenerator steps=500, FiniteStateRed + EnterpriseGreen, numpy PCG64 seeds
1000..1000+N-1 through the reference's own `CybORG(seed=int)` path, i.e. WITHOUT the recording generator) under a
synthetic blue policy and stores, per episode, the team reward of every tick.  tests/test_distribution.py compares the
engine's production RNG mode (Philox, a different stream by design) against these samples: same distribution of episode
returns, not the same trajectories.

    python oracle/ref_returns.py --policy sleep --episodes 192 --procs 8     -> tests/golden/ref_rewards_sleep.npz
    python oracle/ref_returns.py --policy random --episodes 192 --procs 8    -> tests/golden/ref_rewards_random.npz
"""
from __future__ import annotations

import argparse
import multiprocessing as mp
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness as H  # noqa: E402


def one(args):
    seed, policy, steps = args
    H._ensure_path()
    from CybORG import CybORG
    from CybORG.Agents import SleepAgent, EnterpriseGreenAgent, FiniteStateRedAgent
    from CybORG.Agents.Wrappers import BlueFlatWrapper
    from CybORG.Simulator.Scenarios import EnterpriseScenarioGenerator
    sg = EnterpriseScenarioGenerator(blue_agent_class=SleepAgent, green_agent_class=EnterpriseGreenAgent,
                                     red_agent_class=FiniteStateRedAgent, steps=steps)
    env = BlueFlatWrapper(CybORG(sg, seed=seed))
    env.reset()
    prng = np.random.Generator(np.random.PCG64(seed + 7919))
    rew = np.zeros(steps - 1, dtype=np.float32)
    for t in range(steps - 1):
        if policy == "sleep":
            a = H.SLEEP_INDEX
        else:   # uniform over the whole index range, masked-out entries act as Sleep (bench.py / cc4o_run_episodes policy)
            a = [int(prng.integers(0, n)) for n in H.N_ACTIONS]
        _, r, term, _, _ = env.step({f"blue_agent_{i}": int(x) for i, x in enumerate(a)})
        rew[t] = r["blue_agent_0"]
    return rew


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--policy", default="sleep", choices=["sleep", "random"])
    ap.add_argument("--episodes", type=int, default=192)
    ap.add_argument("--procs", type=int, default=os.cpu_count())
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--first-seed", type=int, default=1000)
    ap.add_argument("--name", default=None, help="output file stem (default ref_rewards_<policy>)")
    ap.add_argument("--out", default=os.path.join(HERE, "..", "tests", "golden"))
    a = ap.parse_args()
    t0 = time.time()
    with mp.Pool(a.procs) as pool:
        rows = pool.map(one, [(a.first_seed + i, a.policy, a.steps) for i in range(a.episodes)], chunksize=1)
    rew = np.stack(rows)
    path = os.path.join(a.out, f"{a.name or 'ref_rewards_' + a.policy}.npz")
    np.savez_compressed(path, reward=rew.astype(np.float16), policy=np.array(a.policy), steps=np.int32(a.steps), first_seed=np.int64(a.first_seed))
    tot = rew.sum(1)
    print(f"{path}: {a.episodes} episodes, return mean {tot.mean():.1f} sd {tot.std(ddof=1):.1f}, {time.time() - t0:.0f}s")
