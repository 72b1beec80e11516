"""Parity sweep (TEST INFRASTRUCTURE, build container only): run N fresh episodes of the unmodified reference through
the recording generator and replay each one in the C oracle, comparing the canonical state, the five observation
vectors, reward and done after every tick.  Nothing is written to the repository; the command prints one line per
episode and a summary.  Used to widen the evidence behind the committed golden fixtures.

    python oracle/replay_sweep.py --episodes 70 --procs 7
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

POLICIES = ["random", "masked", "reactive", "lockdown", "sleep", "decoy"]


def one(job):
    seed, policy, steps = job
    import canon
    import oracle_py as O
    import ref_harness as H
    fx = H.run_episode(seed, policy, steps)
    env = O.OracleEnv(steps=steps, tape=O.tape_from_fixture(fx))
    env.reset()
    off = fx["canon_off"]
    def same_state(t):
        exp, got = fx["canon"][off[t]:off[t + 1]], env.export()
        return len(exp) == len(got) and not np.any(exp != got)
    if not same_state(0) or not np.array_equal(env.observe(), fx["obs"][0]) or not np.array_equal(env.mask(), fx["mask"]):
        return seed, policy, steps, "reset differs"
    for t in range(len(fx["actions"])):
        rc, obs, rew, done = env.step(fx["actions"][t], fx["msgs"][t])
        if env.error()[0]:
            return seed, policy, steps, f"tick {t}: oracle error {env.error()}"
        if not same_state(t + 1):
            exp, got = fx["canon"][off[t + 1]:off[t + 2]], env.export()
            return seed, policy, steps, f"tick {t}: " + "; ".join(canon.diff(canon.decode(exp), canon.decode(got))[:4])
        if not np.array_equal(obs, fx["obs"][t + 1]) or rew != fx["reward"][t] or done != fx["done"][t]:
            return seed, policy, steps, f"tick {t}: outputs differ"
    return seed, policy, steps, None


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--episodes", type=int, default=70)
    ap.add_argument("--procs", type=int, default=os.cpu_count())
    ap.add_argument("--first-seed", type=int, default=100)
    a = ap.parse_args()
    jobs = [(a.first_seed + i, POLICIES[i % len(POLICIES)], (500, 500, 300, 167, 500, 120)[i % 6]) for i in range(a.episodes)]
    t0 = time.time()
    bad = 0
    with mp.Pool(a.procs) as pool:
        for seed, policy, steps, err in pool.imap_unordered(one, jobs, chunksize=1):
            print(f"seed {seed} {policy:8s} steps {steps}: {'ok' if err is None else 'MISMATCH ' + err}", flush=True)
            bad += err is not None
    print(f"{len(jobs) - bad} of {len(jobs)} reference episodes replayed bit-exactly by the oracle in {time.time() - t0:.0f}s")
