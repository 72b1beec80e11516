"""The C oracle against the golden trajectories recorded from the live reference (oracle/ref_harness.py).

Every fixture holds the reference's RNG tape, the blue inputs, and after reset and after every tick the canonical
state (oracle/canon.py), the five observation vectors, the team reward and the done flag.  The oracle must reproduce
all of it bit-for-bit.
"""
import os

import numpy as np
import pytest

import canon
import oracle_py as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FILES = sorted(f for f in os.listdir(GOLDEN) if f.startswith("ep_") and f.endswith(".npz"))


@pytest.mark.parametrize("name", FILES)
def test_oracle_replays_reference(name):
    fx = np.load(os.path.join(GOLDEN, name))
    env = O.OracleEnv(steps=int(fx["steps"]), tape=O.tape_from_fixture(fx), red=str(fx["red"]) if "red" in fx else "fsm")
    env.reset()
    off = fx["canon_off"]

    def check_state(t):
        exp, got = fx["canon"][off[t]:off[t + 1]], env.export()
        if len(exp) != len(got) or np.any(exp != got):
            pytest.fail(f"{name} after tick {t - 1}: " + "; ".join(canon.diff(canon.decode(exp), canon.decode(got))[:8]))

    assert env.error() == (0, "")
    check_state(0)
    assert np.array_equal(env.observe(), fx["obs"][0])
    assert np.array_equal(env.mask(), fx["mask"])
    for t in range(len(fx["actions"])):
        rc, obs, rew, done = env.step(fx["actions"][t], fx["msgs"][t])
        assert env.error() == (0, ""), f"tick {t}"
        check_state(t + 1)
        assert np.array_equal(obs, fx["obs"][t + 1]), f"tick {t}: observation"
        assert rew == fx["reward"][t], f"tick {t}: reward"
        assert done == fx["done"][t], f"tick {t}: done"
    assert fx["done"][-1] == (len(fx["actions"]) == int(fx["steps"]) - 1)


PX_FILES = sorted(f for f in os.listdir(GOLDEN) if f.startswith("px_") and f.endswith(".npz"))


@pytest.mark.parametrize("name", PX_FILES)
def test_oracle_in_production_mode_equals_reference_on_the_philox_seam(name):
    """PRODUCTION-mode parity.  tests/golden/px_*.npz are episodes of the unmodified reference whose random source was
    ref_harness.PhiloxSeamGenerator: every modelled draw site received the word the engine's Philox stream of
    (seed, episode, actor, tick) holds for it.  The oracle in rng_mode 0 (Philox, no tape) with the same (seed, episode)
    must then BE that episode: canonical state, observations, reward and done after every tick."""
    fx = np.load(os.path.join(GOLDEN, name))
    assert str(fx["rng"]) == "philox"
    env = O.OracleEnv(seed=int(fx["seed"]), episode=int(fx["episode"]), steps=int(fx["steps"]), red=str(fx["red"]))
    env.reset()
    off = fx["canon_off"]

    def check_state(t):
        exp, got = fx["canon"][off[t]:off[t + 1]], env.export()
        if len(exp) != len(got) or np.any(exp != got):
            pytest.fail(f"{name} after tick {t - 1}: " + "; ".join(canon.diff(canon.decode(exp), canon.decode(got))[:8]))

    check_state(0)
    assert np.array_equal(env.observe(), fx["obs"][0])
    assert np.array_equal(env.mask(), fx["mask"])
    for t in range(len(fx["actions"])):
        rc, obs, rew, done = env.step(fx["actions"][t], fx["msgs"][t])
        assert env.error() == (0, ""), f"tick {t}"
        check_state(t + 1)
        assert np.array_equal(obs, fx["obs"][t + 1]), f"tick {t}: observation"
        assert rew == fx["reward"][t], f"tick {t}: reward"
        assert done == fx["done"][t], f"tick {t}: done"


def test_oracle_detects_a_corrupted_tape():
    fx = dict(np.load(os.path.join(GOLDEN, FILES[-1])))
    recs, seg = O.tape_from_fixture(fx)
    first_tick = int(seg[1])
    recs["site"][first_tick] ^= 1     # first draw of tick 0
    env = O.OracleEnv(steps=int(fx["steps"]), tape=(recs, seg))
    env.reset()
    env.step(fx["actions"][0], fx["msgs"][0])
    assert env.error()[0] & 1


def test_reward_table_matches_reference_tables():
    """The 81 values of BlueRewardMachine.get_phase_rewards (BlueRewardMachine.py:35-65), read from the LIVE reference into
    tests/golden/ref_tables.json (oracle/ref_tables.py calls the reference's method), against what the oracle charges.  The
    oracle counts its reward events per (mission phase, kind, subnet); on every tick of 240 production-mode episodes the
    reward must equal sum(new events x the reference's value) + the action cost (-1 per Restore submitted, Restore.py:70),
    and every cell of the eight populated subnets that carries a non-zero value must have been exercised."""
    import json
    tab = json.load(open(os.path.join(GOLDEN, "ref_tables.json")))
    ref = np.zeros((3, 3, 9))
    for p in range(3):
        for sn, v in tab["reward"][str(p)].items():
            ref[p, :, canon.SUBNET_IDX[sn]] = v
    assert tab["cost"]["Restore"] == -1.0 and all(v == 0.0 for k, v in tab["cost"].items() if k != "Restore")
    hit = np.zeros((3, 3, 9), dtype=np.int64)
    g = np.random.default_rng(5)
    for ep in range(240):
        env = O.OracleEnv(seed=31, episode=ep, steps=500)
        env.reset()
        last = env.reward_breakdown().astype(np.int64)
        mask = env.mask()
        for t in range(499):
            if ep % 3 == 0:
                a = [49, 49, 49, 49, 145]                                     # all Sleep: red reaches the operational zones
            elif ep % 3 == 1:
                a = [int(g.integers(0, k)) for k in (82, 82, 82, 82, 242)]    # uniform indices: Restores, Blocks
            else:                                                             # traffic control only: blocked green access (ASF) everywhere
                a = [int(g.integers(50, 66)) for _ in range(4)] + [int(g.integers(146, 194))]
            rc, _, rew, _ = env.step(a)
            now = env.reward_breakdown().astype(np.int64)
            d = now - last
            last = now
            restores = sum(1 for b, i in enumerate(a) if (2 * (48 if b == 4 else 16) + 1 <= i < 3 * (48 if b == 4 else 16) + 1) and mask[b, i])
            assert rew == float((d * ref).sum()) - restores, f"episode {ep} tick {t}: reward {rew}, events {np.argwhere(d)}"
            hit += d
        assert env.error()[0] == 0
    missing = [(p, "LWF ASF RIA".split()[k], canon.SUBNETS_CREATION[s]) for p in range(3) for k in range(3) for s in range(8)
               if ref[p, k, s] != 0 and hit[p, k, s] == 0]
    assert not missing, f"reward cells never exercised: {missing}"
    assert hit[:, :, 8].sum() == 0          # nothing ever happens in the internet subnet (no green agents, no OT service, no red target)


def test_philox_mode_is_deterministic_and_sharding_invariant():
    a = O.OracleEnv(seed=7, episode=3, steps=100)
    b = O.OracleEnv(seed=7, episode=3, steps=100)
    a.reset(), b.reset()
    assert np.array_equal(a.export(), b.export())
    rng = np.random.default_rng(0)
    for _ in range(40):
        act = [int(rng.integers(0, n)) for n in (82, 82, 82, 82, 242)]
        ra, rb = a.step(act), b.step(act)
        assert ra[2] == rb[2] and np.array_equal(ra[1], rb[1])
    assert np.array_equal(a.export(), b.export())
    assert a.error()[0] == 0
    c = O.OracleEnv(seed=7, episode=4, steps=100)
    c.reset()
    assert not np.array_equal(a.export()[:400], c.export()[:400])


def test_mission_phase_split():
    """EnterpriseScenarioGenerator._generate_mission_phases :854-860 (reference test_mission_phase.py:95-101)."""
    for steps, changes in ((9, (3, 6)), (10, (4, 7)), (11, (4, 8)), (100, (34, 67)), (500, (167, 334))):
        env = O.OracleEnv(seed=1, episode=0, steps=steps)
        env.reset()
        phases = []
        for t in range(steps - 1):
            env.step([49, 49, 49, 49, 145])
            phases.append(int(env.export()[0]))
        got = tuple(i for i in range(1, len(phases)) if phases[i] != phases[i - 1])
        assert got == tuple(c for c in changes if c < steps - 1), (steps, got)


def test_episode_length_and_done():
    env = O.OracleEnv(seed=2, episode=0, steps=30)
    env.reset()
    dones = [env.step([49, 49, 49, 49, 145])[3] for _ in range(29)]
    assert dones == [0] * 28 + [1]      # EnterpriseScenarioGenerator.determine_done :862-871
