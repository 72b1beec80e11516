"""Is the recording generator's behaviour the reference's behaviour?  (TEST INFRASTRUCTURE; needs /root/reference.)

The golden tapes are recorded with ref_harness.TapeGenerator, a duck-typed generator that returns Python ints / floats /
the list elements themselves, while a user's `CybORG(seed=int)` runs on numpy.random.Generator, whose `choice` returns
numpy scalars and converts sequences to arrays first.  If any code path of the reference depended on those types, tape
replay (and the Philox seam) would pin a behaviour users never see.  This script closes that gap: TapeGenerator consumes
the underlying PCG64 exactly as numpy's own methods do (Generator.choice(a) == a[integers(0, len(a))], choice(p=) ==
searchsorted(cdf, random()), integers(lo, hi) == lo + integers(0, hi - lo)); the only method that consumes the stream
differently is shuffle, so the check uses numpy's own shuffle.  Then `CybORG(seed=s)` and `CybORG(seed=recorder(s))` must
produce IDENTICAL reward / observation trajectories.  They do (run log in DESIGN.md section 2).

    python oracle/seam_semantics_check.py --seeds 5 --ticks 499
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness as H  # noqa: E402


class SameStreamRecorder(H.TapeGenerator):
    def shuffle(self, x):
        site, actor = self._context()
        self._g.shuffle(x)
        self._log(site, actor, H.KIND_INT, len(x), 0)


def run(seed, policy, ticks, recorder: bool):
    H._ensure_path()
    from CybORG import CybORG
    from CybORG.Agents import SleepAgent, EnterpriseGreenAgent, FiniteStateRedAgent
    from CybORG.Agents.Wrappers import BlueFlatWrapper
    from CybORG.Simulator.Scenarios import EnterpriseScenarioGenerator
    sg = EnterpriseScenarioGenerator(blue_agent_class=SleepAgent, green_agent_class=EnterpriseGreenAgent,
                                     red_agent_class=FiniteStateRedAgent, steps=500)
    env = BlueFlatWrapper(CybORG(sg, seed=SameStreamRecorder(seed) if recorder else seed))
    obs, _ = env.reset()
    prng = np.random.Generator(np.random.PCG64(seed + 7919))
    mask = np.zeros((5, 242), dtype=np.uint8)
    for i in range(5):
        m = env.action_mask(f"blue_agent_{i}")
        mask[i, :len(m)] = m
    out = []
    for _ in range(ticks):
        o = np.zeros((5, 210), dtype=np.uint8)
        for i in range(5):
            v = np.asarray(obs[f"blue_agent_{i}"])
            o[i, :len(v)] = v
        a, m = H.blue_policy(policy, prng, mask, o)
        obs, r, term, trunc, _ = env.step({f"blue_agent_{i}": int(x) for i, x in enumerate(a)},
                                          messages={f"blue_agent_{i}": np.array([(int(v) >> b) & 1 for b in range(8)], dtype=bool) for i, v in enumerate(m)})
        out.append((float(r["blue_agent_0"]), tuple(int(x) for k in sorted(obs) for x in obs[k])))
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, default=3)
    ap.add_argument("--ticks", type=int, default=499)
    a = ap.parse_args()
    bad = 0
    for k in range(a.seeds):
        for policy in ("sleep", "random", "reactive"):
            seed = 9100 + k
            x, y = run(seed, policy, a.ticks, False), run(seed, policy, a.ticks, True)
            same = x == y
            bad += not same
            print(f"seed {seed} {policy}: numpy Generator vs recording generator over {a.ticks} ticks: "
                  f"{'IDENTICAL' if same else 'DIFFERENT'} (return {sum(r for r, _ in x):.0f})", flush=True)
    sys.exit(1 if bad else 0)
