"""Distributional parity of the production RNG mode against the live reference (tertiary parity).

What makes production mode exact is elsewhere: tests/golden/px_*.npz are episodes of the unmodified reference driven by the
engine's own Philox words (ref_harness.PhiloxSeamGenerator), reproduced bit for bit by the oracle and the CUDA engine in
production mode; oracle/seam_semantics_check.py shows that the recording generator behaves exactly like numpy's Generator on
the same stream.  Given that, the engine's outcomes are the reference's outcomes under a different but equally uniform stream,
and this test only guards the remaining assumption (the words are uniform and independent) the statistical way.

tests/golden/ref_rewards_{sleep,random}.npz hold the per-tick team reward of full episodes of the unmodified reference under
numpy PCG64 (oracle/ref_returns.py): 192 episodes under uniform-random blue, 2 072 under all-Sleep blue (five runs, seeds
listed in the file).  The oracle (bit-exact with the CUDA engine in this mode) runs 2 500 other episodes; compared are the
episode return overall and per mission phase (Welch z, gate |z| < 3), the spread, and the shape (two-sample KS).

History of the all-Sleep comparison: round 1 had 872 reference episodes (-6 831 +- 47) against 768 engine episodes (-6 622 +- 49),
z = -3.1 overall and -4.2 in mission phase 1.  Both samples turned out to sit in the tails: 1 200 further reference episodes
give -6 670 +- 40, and 9 000 further engine episodes (seeds 1234 and 77) give -6 770 +- 18 / -6 756 (phase 1: -2 392 +- 10 against
the reference's pooled -2 383 +- 18, z = +0.4).  Pooled reference (2 072) against the 6 000-episode engine run: z = +0.95.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle_py as O  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
N_ENGINE = 2500
PHASES = [(0, 167), (167, 334), (334, 499)]     # EnterpriseScenarioGenerator._generate_mission_phases :854-860 (500 -> 167/167/166)


def _engine_rewards(policy, n):
    """per-tick rewards [n, 499] of the oracle in production (Philox) mode"""
    out = np.zeros((n, 499), dtype=np.float32)
    g = np.random.default_rng(12345)
    for i in range(n):
        env = O.OracleEnv(seed=1234, episode=i, steps=500)
        env.reset()
        for t in range(499):
            a = (49, 49, 49, 49, 145) if policy == "sleep" else [int(g.integers(0, k)) for k in (82, 82, 82, 82, 242)]
            rc, _, r, d = env.step(np.array(a, dtype=np.int32))
            out[i, t] = r
        assert d == 1 and env.error()[0] == 0
    return out


def _ks(a, b):
    """two-sample Kolmogorov-Smirnov statistic and its asymptotic p-value"""
    a, b = np.sort(a), np.sort(b)
    allv = np.concatenate([a, b])
    d = np.max(np.abs(np.searchsorted(a, allv, side="right") / len(a) - np.searchsorted(b, allv, side="right") / len(b)))
    en = np.sqrt(len(a) * len(b) / (len(a) + len(b)))
    lam = (en + 0.12 + 0.11 / en) * d
    p = 2 * sum((-1) ** (k - 1) * np.exp(-2 * k * k * lam * lam) for k in range(1, 101))
    return d, float(min(max(p, 0.0), 1.0))


@pytest.mark.parametrize("policy", ["sleep", "random"])
def test_episode_return_distribution_matches_reference(policy):
    path = os.path.join(GOLDEN, f"ref_rewards_{policy}.npz")
    ref = np.load(path)["reward"].astype(np.float64)
    assert ref.shape[1] == 499 and ref.shape[0] >= 100
    eng = _engine_rewards(policy, N_ENGINE).astype(np.float64)
    report = []
    for name, (lo, hi) in [("episode", (0, 499))] + [(f"phase {i}", p) for i, p in enumerate(PHASES)]:
        r, e = ref[:, lo:hi].sum(1), eng[:, lo:hi].sum(1)
        z = (r.mean() - e.mean()) / np.sqrt(r.var(ddof=1) / len(r) + e.var(ddof=1) / len(e))
        report.append(f"{name}: reference {r.mean():.1f} +- {r.std(ddof=1):.1f}, engine {e.mean():.1f} +- {e.std(ddof=1):.1f}, z = {z:.2f}")
        assert abs(z) < 3.0, "; ".join(report)
        # the spread must agree as well (ratio of standard deviations; 192 samples give ~ +-10 % at 2 sigma)
        assert 0.75 < r.std(ddof=1) / e.std(ddof=1) < 1.33, "; ".join(report)
    d, p = _ks(ref.sum(1), eng.sum(1))
    report.append(f"KS D = {d:.3f}, p = {p:.3f}")
    print("\n".join(report))
    assert p > 1e-3, "; ".join(report)
