"""The mirrored wrapper API, checked the way the reference checks its own wrappers
(CybORG/Tests/test_cc4/test_BlueEnterpriseWrapper.py, test_mission_phase.py, test_cc4_seed.py)."""
import json
import os
import re

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from cage_challenge_4_b200.wrappers import (BlueEnterpriseWrapper, BlueFlatWrapper, CybORG, EnterpriseGreenAgent,  # noqa: E402
                                            EnterpriseMAE, EnterpriseScenarioGenerator, FiniteStateRedAgent, SleepAgent,
                                            SUBNETS_SORTED)

NUM_SUBNETS, MAX_HOSTS, MESSAGE_LENGTH, NUM_MESSAGES = 9, 16, 8, 4
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def make(seed=123, steps=100, cls=BlueEnterpriseWrapper, **kw):
    sg = EnterpriseScenarioGenerator(blue_agent_class=SleepAgent, green_agent_class=EnterpriseGreenAgent,
                                     red_agent_class=FiniteStateRedAgent, steps=steps)
    return cls(CybORG(sg, seed=seed), **kw)


def test_space_sizes_and_action_counts():
    """test_BlueEnterpriseWrapper.py:30-43 (slice layout) and :86-98 (action count formula)."""
    env = make()
    obs, info = env.reset()
    for i, a in enumerate(env.possible_agents):
        n_sub = 3 if i == 4 else 1
        assert len(obs[a]) == 1 + n_sub * (3 * NUM_SUBNETS + 2 * MAX_HOSTS) + NUM_MESSAGES * MESSAGE_LENGTH
        assert env.observation_space(a).contains(obs[a])
        n_actions = 2 + n_sub * (3 * MAX_HOSTS) + n_sub * 2 * (NUM_SUBNETS - 1) + n_sub * MAX_HOSTS
        assert env.action_space(a).n == n_actions == len(info[a]["action_mask"]) == len(env.action_labels(a))
        assert obs[a].dtype == np.int64


def test_observation_slices():
    env = make()
    obs, _ = env.reset()
    for i, a in enumerate(env.possible_agents):
        v = obs[a]
        assert v[0] == 0                                         # mission phase
        for k, sn in enumerate(env.subnets(a)):
            base = 1 + k * 59
            assert v[base:base + 9].tolist() == [int(s == sn) for s in SUBNETS_SORTED]   # subnet one-hot
            assert v[base + 9:base + 18].sum() == 0                                        # nothing blocked at reset
            assert v[base + 27:base + 59].sum() == 0                                       # no alerts at reset
        assert v[-32:].sum() == 0


def test_comms_policy_matrices():
    """Golden adjacency of test_BlueEnterpriseWrapper.py:262-296 / README.md:107-159: bit = 'should be blocked'."""
    order = SUBNETS_SORTED
    core = ["internet_subnet", "admin_network_subnet", "office_network_subnet", "public_access_zone_subnet",
            "contractor_network_subnet", "restricted_zone_a_subnet", "restricted_zone_b_subnet"]
    def expected(phase, sn):
        edges = {frozenset((a, b)) for a in core for b in core if a != b}
        edges |= {frozenset(("restricted_zone_a_subnet", "operational_zone_a_subnet")), frozenset(("restricted_zone_b_subnet", "operational_zone_b_subnet"))}
        if phase == 1:
            edges -= {frozenset(("restricted_zone_a_subnet", x)) for x in ("operational_zone_a_subnet", "contractor_network_subnet", "restricted_zone_b_subnet", "internet_subnet")}
        if phase == 2:
            edges -= {frozenset(("restricted_zone_b_subnet", x)) for x in ("operational_zone_b_subnet", "contractor_network_subnet", "restricted_zone_a_subnet", "internet_subnet")}
        return [int(frozenset((sn, o)) not in edges) for o in order]
    env = make(steps=12)            # phases of 4 ticks
    obs, _ = env.reset()
    seen = set()
    for t in range(11):
        for i, a in enumerate(env.possible_agents):
            phase = int(obs[a][0])
            seen.add(phase)
            for k, sn in enumerate(env.subnets(a)):
                assert obs[a][1 + k * 59 + 18:1 + k * 59 + 27].tolist() == expected(phase, sn), (a, sn, phase)
        if not env.agents:
            break
        obs, *_ = env.step({a: 49 if a != "blue_agent_4" else 145 for a in env.agents})
    assert seen == {0, 1, 2}


def test_block_and_allow_show_up_in_the_vector():
    env = make()
    env.reset()
    a = "blue_agent_0"
    labels = env.action_labels(a)
    strip = [re.sub(r" \([0-9./]+\)", "", l) for l in labels]        # labels carry the CIDRs like the reference's (BlueFixedActionWrapper.py:283)
    blk = strip.index("BlockTrafficZone restricted_zone_a_subnet <- contractor_network_subnet")
    alw = strip.index("AllowTrafficZone restricted_zone_a_subnet <- contractor_network_subnet")
    assert labels[blk] == "BlockTrafficZone restricted_zone_a_subnet (10.0.0.0/24) <- contractor_network_subnet (10.0.4.0/24)"
    obs, *_ = env.step({a: blk})
    pos = 1 + 9 + SUBNETS_SORTED.index("contractor_network_subnet")
    assert obs[a][pos] == 1 and obs[a][1 + 9:1 + 18].sum() == 1
    obs, *_ = env.step({a: alw})
    assert obs[a][1 + 9:1 + 18].sum() == 0


def test_messages_are_delivered_in_sender_order():
    """test_BlueEnterpriseWrapper.py:439-462."""
    env = make()
    env.reset()
    msgs = {f"blue_agent_{i}": np.array([(i + 1) >> b & 1 for b in range(8)], dtype=bool) for i in range(5)}
    obs, *_ = env.step({}, messages=msgs)
    for i, a in enumerate(env.possible_agents):
        others = [j for j in range(5) if j != i]
        got = obs[a][-32:].reshape(4, 8)
        for k, j in enumerate(others):
            assert got[k].tolist() == msgs[f"blue_agent_{j}"].astype(int).tolist()
    with pytest.raises(AssertionError):
        env.step({}, messages={"blue_agent_0": np.zeros(5, dtype=bool)})


def test_mask_matches_hosts_and_invalid_actions_sleep():
    env = make()
    _, info = env.reset()
    for a in env.possible_agents:
        labels, mask = env.action_labels(a), info[a]["action_mask"]
        assert all(m == (not l.startswith("[Invalid]")) for l, m in zip(labels, mask))
        assert 34 <= sum(mask)          # Monitor + Sleep + 16 zone actions + 4 commands x (3 users + 1 server)


def test_same_seed_same_trajectory_and_reset_gives_new_scenario():
    """test_cc4_seed.py:16-31 (self-consistency)."""
    a, b = make(seed=7), make(seed=7)
    oa, _ = a.reset()
    ob, _ = b.reset()
    rng = np.random.default_rng(0)
    for _ in range(20):
        acts = {f"blue_agent_{i}": int(rng.integers(0, 82)) for i in range(5)}
        ra, rb = a.step(acts), b.step(acts)
        assert all(np.array_equal(ra[0][k], rb[0][k]) for k in ra[0]) and ra[1] == rb[1]
    m1 = a.reset()[1]["blue_agent_4"]["action_mask"]
    m2 = a.reset()[1]["blue_agent_4"]["action_mask"]
    m3 = a.reset()[1]["blue_agent_4"]["action_mask"]
    assert not (m1 == m2 == m3)


def test_episode_ends_and_agents_list_empties():
    env = make(steps=10, cls=EnterpriseMAE)
    env.reset()
    for t in range(9):
        obs, rew, term, trunc, info = env.step({"actions": {a: 0 for a in env.agents}})
        assert term["__all__"] is False
        assert trunc["__all__"] == (t == 8)
    assert env.agents == []
    with pytest.raises(NotImplementedError):
        EnterpriseScenarioGenerator(red_agent_class=SleepAgent)
    from cage_challenge_4_b200.wrappers import DiscoveryFSRed, RandomSelectRedAgent
    for cls_ in (DiscoveryFSRed, RandomSelectRedAgent):        # the other scripted red agents run through the same facade
        e2 = BlueFlatWrapper(CybORG(EnterpriseScenarioGenerator(red_agent_class=cls_, steps=30), seed=5))
        e2.reset()
        for _ in range(29):
            _, rew, term, *_ = e2.step({})
        assert term["blue_agent_0"] is True and e2.env.engine.error_flags()[0] == 0


def test_pad_spaces_against_the_reference():
    """BlueFlatWrapper(pad_spaces=True) (BlueFixedActionWrapper.py:311-328, BlueFlatWrapper.py:251-254) against the live
    reference on the Philox seam (tests/golden/pad_*.npz, oracle/ref_harness.py --padded): spaces, masks, labels (CIDRs aside),
    and every padded observation vector of 99 ticks of indices drawn over the whole padded range."""
    fx = np.load(os.path.join(GOLDEN, "pad_s35_100_99.npz"))
    env = make(seed=int(fx["seed"]), steps=int(fx["steps"]), cls=BlueFlatWrapper, pad_spaces=True)
    obs, info = env.reset(seed=int(fx["seed"]))         # re-seeding restarts the episode numbering: (seed, episode 0), the fixture's episode
    ref_labels = json.loads(str(fx["labels"]))
    for b, a in enumerate(env.possible_agents):
        assert [env.action_space(a).n, len(env.observation_space(a))] == fx["spaces"][b].tolist() == [242, 210]
        assert info[a]["action_mask"] == fx["mask"][b].astype(bool).tolist() == env.action_mask(a)
        strip = lambda ls: [re.sub(r" \([0-9./]+\)", "", l) for l in ls]
        assert strip(env.action_labels(a)) == strip(ref_labels[b])
        assert np.array_equal(obs[a], fx["obs"][0][b]) and obs[a].dtype == np.int64
    for t in range(len(fx["actions"])):
        obs, rew, term, trunc, info = env.step({a: int(fx["actions"][t][b]) for b, a in enumerate(env.possible_agents)})
        for b, a in enumerate(env.possible_agents):
            assert np.array_equal(obs[a], fx["obs"][t + 1][b]), (t, a)
        assert rew["blue_agent_0"] == fx["reward"][t]
    with pytest.raises(IndexError):
        env.step({"blue_agent_0": 242})


def test_evaluation_protocol_runs_unchanged_loop():
    """The reference's evaluation loop (evaluation.py:89-124) over the mirrored API, plus the batched variant."""
    from cage_challenge_4_b200.evaluation import MLPPolicy, run_evaluation, run_evaluation_batched

    class DummyAgent:                                   # example_submission/dummy_agent.py: last action index
        def get_action(self, observation, action_space):
            return action_space.n - 1

    class Submission:
        NAME, TEAM, TECHNIQUE = "n", "t", "q"
        AGENTS = {f"blue_agent_{i}": DummyAgent() for i in range(5)}

        @staticmethod
        def wrap(env):
            return EnterpriseMAE(env)

    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        out = run_evaluation(Submission, log_path=tmp, max_eps=2, seed=3, write_to_file=True)
        assert out["episodes"] == 2 and out["mean"] < 0
        # the five log files of evaluation.py:141-196
        assert sorted(os.listdir(tmp)) == ["actions.txt", "full.txt", "scores.txt", "summary.json", "summary.txt"]
        summary = json.load(open(os.path.join(tmp, "summary.json")))
        assert summary["reward"]["mean"] == out["mean"] and summary["parameters"] == {"seed": 3, "episode_length": 500, "max_episodes": 2}
        full = open(os.path.join(tmp, "full.txt")).read()
        assert "DeployDecoy" in full and "observations:" in full and f"total reward: {out['rewards'][0]}" in full


def test_batched_evaluation_totals_equal_an_oracle_replay_and_do_not_depend_on_sharding():
    """run_evaluation_batched(MLPPolicy): the per-episode totals must be what the C oracle computes when it is fed the very
    actions the policy emitted (evaluation.py:104-110 bookkeeping: the done tick's reward is dropped), with and without the
    CUDA graph; and two shards (the ranks of a 2-GPU run, here run one after the other) must give the one-shard totals."""
    import oracle_py as O
    from cage_challenge_4_b200.evaluation import MLPPolicy, run_evaluation_batched
    pol = MLPPolicy(seed=4).cuda()
    n, steps, seed = 12, 120, 5
    eager = run_evaluation_batched(pol, max_eps=n, seed=seed, steps=steps, graph=False, return_actions=True)
    acts = eager["actions"].cpu().numpy()
    assert acts.shape == (steps - 1, n, 5)
    for i in range(n):
        env = O.OracleEnv(seed=seed, episode=i, steps=steps)
        env.reset()
        total = 0.0
        for t in range(steps - 1):
            rc, _, r, d = env.step(acts[t, i])
            assert d == (t == steps - 2)
            if not d:
                total += r
        assert total == eager["rewards"][i], (i, total, eager["rewards"][i])
    graphed = run_evaluation_batched(pol, max_eps=n, seed=seed, steps=steps, graph=True)
    assert graphed["rewards"] == eager["rewards"] and graphed["stdev"] > 0
    lo = run_evaluation_batched(pol, max_eps=n, seed=seed, steps=steps, rank=0, world=2)
    hi = run_evaluation_batched(pol, max_eps=n, seed=seed, steps=steps, rank=1, world=2)
    # world=2 without a process group: aggregate_scores is exercised by tests/test_sharding_cpu.py; here the shards' own totals
    from cage_challenge_4_b200.evaluation import shard
    assert shard(n, 0, 2) == (0, 6) and shard(n, 1, 2) == (6, 12)
    assert lo["local_rewards"] + hi["local_rewards"] == eager["rewards"]


def test_random_blue_policy_samples_class_then_parameter():
    """cc4BlueRandomAgent (RandomAgent.py:28-56, 66-69) on device: the action CLASS is uniform over Sleep / Monitor / Analyse /
    Remove / DeployDecoy / Restore (Block / Allow excluded), then the host is uniform over the existing ones."""
    import torch
    from cage_challenge_4_b200.engine import BatchedCC4
    from cage_challenge_4_b200.evaluation import RandomBluePolicy, run_evaluation_batched
    env = BatchedCC4(4096, device=0, seed=5)
    obs, mask = env.reset()
    pol = RandomBluePolicy(seed=1)
    counts = np.zeros((2, 6))
    for _ in range(10):
        a = pol(obs, mask)
        assert bool(torch.gather(mask, 2, a.long().unsqueeze(-1)).all())
        for k, (agents, H, T) in enumerate(((slice(0, 4), 16, 8), (slice(4, 5), 48, 24))):
            x = a[:, agents].cpu().numpy().ravel()
            assert not ((x >= 3 * H + 2) & (x < 3 * H + 2 + 2 * T)).any()          # no Allow / Block
            cls = np.select([x == 3 * H + 1, x == H, x < H, x < 2 * H + 1, x < 3 * H + 1], [0, 1, 2, 3, 5], default=4)
            counts[k] += np.bincount(cls, minlength=6)
        obs, rew, done = env.step(a)
    freq = counts / counts.sum(1, keepdims=True)
    assert np.abs(freq - 1 / 6).max() < 0.01, freq              # 163 840 / 40 960 draws: three sigma is 0.003 / 0.006
    # hosts uniform over the existing ones: every valid Analyse index of agent 0 in episode 0 comes up
    valid = torch.nonzero(mask[0, 0, :16]).ravel().tolist()
    seen = set()
    for _ in range(300):
        x = int(pol(obs, mask)[0, 0])
        if x < 16:
            seen.add(x)
    assert seen == set(valid)
    out = run_evaluation_batched(RandomBluePolicy(seed=2), max_eps=32, seed=3, steps=60)
    assert out["episodes"] == 32 and out["mean"] < 0


def test_last_action_is_the_index_and_prints_as_the_label():
    """env.py:300 get_last_action feeds the evaluation logs (evaluation.py:111-117): the mirrored call returns the submitted
    index, printing as the action's label."""
    env = make()
    env.reset()
    cyborg = env.env
    assert cyborg.get_last_action("blue_agent_0") is None
    labels = env.action_labels("blue_agent_4")
    idx = next(i for i, l in enumerate(labels) if l.startswith("Restore") and "Invalid" not in l)
    env.step({"blue_agent_4": idx})
    last = cyborg.get_last_action("blue_agent_4")
    assert last == [idx] and str(last[0]) == labels[idx] and repr(last) == f"[{labels[idx]}]"
    assert str(cyborg.get_last_action("blue_agent_0")[0]) == "Sleep"


def test_true_state_dict_of_the_wrapped_episode():
    """env.py:265 get_agent_state('True'): hosts of the episode with their processes / services / sessions."""
    env = make()
    env.reset()
    for _ in range(30):
        env.step({})
    st = env.env.get_agent_state("True")
    assert st["step"] == 30 and st["mission_phase"] == 0
    hosts = st["hosts"]
    assert "root_internet_host_0" in hosts and "restricted_zone_a_subnet_router" in hosts
    servers = [h for h in hosts if h.endswith("_server_host_0")]
    assert len(servers) == 8 and all(any(p["process_name"] == "sshd" for p in hosts[h]["Processes"]) for h in servers)
    assert sum(len(r["sessions"]) for r in st["red"].values()) >= 1
    mask_hosts = {h for a in env.possible_agents for h, m in zip([x for x in env.hosts(a) if "router" not in x], env.action_mask(a)) if m}
    assert {f"{h}" for h in mask_hosts} <= set(hosts)        # every host a blue agent may Analyse exists in the true state
    assert set(env.env.get_agent_state("blue_agent_0")) >= {"success", "message"}


def test_facade_surface_of_env_py():
    """CybORG/env.py: parallel_step :95, step :125, get_observation :245, get_action_space :266, get_last_action :300,
    get_ip_map :327, get_rewards :348, get_attr :374, agents :394 — and BlueFixedActionWrapper.actions() :360."""
    from cage_challenge_4_b200.wrappers import ActionDescriptor, BlueFixedActionWrapper, TernaryEnum
    sg = EnterpriseScenarioGenerator(blue_agent_class=SleepAgent, green_agent_class=EnterpriseGreenAgent, red_agent_class=FiniteStateRedAgent, steps=100)
    cy = CybORG(sg, seed=11)
    assert cy.agents == cy.active_agents == [f"blue_agent_{i}" for i in range(5)]
    fixed = BlueFixedActionWrapper(cy)
    acts = fixed.actions("blue_agent_4")
    labels = fixed.action_labels("blue_agent_4")
    assert len(acts) == 242 and all(isinstance(a, ActionDescriptor) for a in acts)
    i_restore = next(i for i, l in enumerate(labels) if l.startswith("Restore"))
    assert str(acts[i_restore]) == labels[i_restore] and acts[i_restore].get_params()["session"] == 0
    assert str(acts[145]) == "Sleep" and str(acts[48]) == "Monitor" and acts[146].name == "AllowTrafficZone"
    assert fixed.get_action_space("blue_agent_4")["actions"] is acts
    # parallel_step with indices: dict observations, team reward dicts, done flags (env.py:95-123)
    obs, rew, done, info = cy.parallel_step({"blue_agent_4": i_restore, "blue_agent_0": 49})
    assert set(obs) == set(cy.agents) and info == {}
    assert obs["blue_agent_4"]["success"] is TernaryEnum.IN_PROGRESS                      # Restore takes 5 ticks (Restore.py:36)
    assert obs["blue_agent_0"]["success"] is TernaryEnum.UNKNOWN and str(obs["blue_agent_0"]["action"]) == "Sleep"
    assert rew["blue_agent_0"]["action_cost"] == -1.0 and set(rew["blue_agent_0"]) == {"BlueRewardMachine", "action_cost"}
    assert cy.get_rewards()["Blue"] == rew["blue_agent_3"] and cy.get_rewards()["Red"] == {"None": 0.0, "action_cost": 0}
    assert done["blue_agent_1"] is False
    for _ in range(4):
        obs, *_ = cy.parallel_step({})
    assert obs["blue_agent_4"]["success"] is TernaryEnum.TRUE and str(obs["blue_agent_4"]["action"]) == labels[i_restore]
    assert str(cy.get_last_action("blue_agent_4")[0]) == "Sleep"
    # descriptors and reference-style Action objects (class name + attributes) are mapped to their index
    class Analyse:                                                                         # what a reference Action object looks like from outside
        def __init__(self, hostname, session=0, agent=None):
            self.hostname, self.session, self.agent = hostname, session, agent
    host = fixed.actions("blue_agent_2")[0].hostname
    assert cy.action_index("blue_agent_2", Analyse(host)) == 0 == cy.action_index("blue_agent_2", fixed.actions("blue_agent_2")[0])
    assert cy.action_index("blue_agent_4", acts[146]) == 146 and cy.action_index("blue_agent_4", acts[145]) == 145
    with pytest.raises(ValueError):
        cy.action_index("blue_agent_2", Analyse("no_such_host"))
    res = cy.step("blue_agent_1", 16)                                                      # CC1/CC2-style call: Monitor
    assert res.observation["success"] is TernaryEnum.TRUE and str(res.observation["action"]) == "Monitor" and res.done is False
    # maps and action space dicts
    ipm, space = cy.get_ip_map(), cy.get_action_space("blue_agent_0")
    assert "root_internet_host_0" in ipm and len(set(ipm.values())) == len(ipm)
    mine = [h for h, known in space["hostname"].items() if known]
    assert mine and all(h.startswith("restricted_zone_a_subnet") for h in mine) and set(space["hostname"]) == set(ipm)
    assert sum(space["subnet"].values()) == 1 and space["session"] == {0: True} and len(space["action"]) == 8
    assert cy.get_attr("seed") == 11 and cy.get_attr("nonexistent") is None and fixed.get_attr("env") is cy
    # the dict-observation wrapper (BlueFixedActionWrapper.reset :76 / step :101)
    o, info = fixed.reset()
    assert set(o["blue_agent_2"]) >= {"success", "message"} and len(info["blue_agent_2"]["action_mask"]) == 82
    o, r, term, trunc, info = fixed.step({"blue_agent_2": 16})
    assert str(o["blue_agent_2"]["action"]) == "Monitor" and isinstance(r["blue_agent_2"], float)


def test_per_episode_wrappers_refuse_a_batch():
    """A wrapper presents ONE episode: on a batch it would advance the other episodes with Sleep and drop their outputs."""
    sg = EnterpriseScenarioGenerator(steps=50)
    cy = CybORG(sg, seed=1, num_envs=4)
    assert cy.engine.num_envs == 4
    for cls in (BlueFlatWrapper, BlueEnterpriseWrapper, EnterpriseMAE):
        with pytest.raises(ValueError, match="ONE episode"):
            cls(cy)
    with pytest.raises(ValueError):
        cy.parallel_step({})
    obs, rew, done = cy.engine.step(torch.tensor([[49, 49, 49, 49, 145]] * 4, dtype=torch.int32, device="cuda"))     # the batch API still works
    assert obs.shape == (4, 5, 210)


def test_dict_observations_against_the_reference():
    """GPU twin of tests/test_state_cpu.py::test_dict_observations_match_the_reference_cpu: the real facade in production
    mode with the fixture's seed IS the reference episode recorded on the Philox seam, so CybORG.get_observation must report
    the reference's success / action / hosts / pids on every tick."""
    for name in sorted(f for f in os.listdir(GOLDEN) if f.startswith("dictobs_")):
        fx = np.load(os.path.join(GOLDEN, name))
        expected = json.loads(str(fx["summaries"]))
        sg = EnterpriseScenarioGenerator(steps=int(fx["steps"]))
        cy = CybORG(sg, seed=int(fx["seed"]))
        cy.enable_dict_observations()
        for t in range(len(fx["actions"])):
            msgs = {f"blue_agent_{b}": np.array([(int(fx["msgs"][t][b]) >> i) & 1 for i in range(8)], dtype=bool) for b in range(5)}
            obs, rew, done, _ = cy.parallel_step({f"blue_agent_{b}": int(fx["actions"][t][b]) for b in range(5)}, messages=msgs)
            assert round(sum(rew["blue_agent_0"].values()), 1) == fx["reward"][t]
            for b in range(5):
                got, exp = obs[f"blue_agent_{b}"], expected[t][b]
                assert got["success"].name == exp["success"], (name, t, b)
                if exp["success"] != "IN_PROGRESS":
                    assert str(got["action"]) == exp["action"], (name, t, b)
                hosts = {k: sorted(p["PID"] for p in v.get("Processes", []) if "PID" in p) for k, v in got.items() if k not in ("success", "action", "message")}
                assert hosts == exp["hosts"], (name, t, b, hosts, exp["hosts"])


def test_the_reference_evaluation_script_runs_unchanged_on_the_compat_package():
    """CybORG/Evaluation/evaluation.py of the UNMODIFIED reference (baseline/_ref, present when the build container installed it)
    with `compat/` first on sys.path: its own run_evaluation + load_submission + example submission drive the B200 engine and
    write the five log files.  Nothing of the reference's simulator is imported: `CybORG` resolves to compat/CybORG."""
    import subprocess
    import sys
    import tempfile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ref_eval = os.path.join(root, "baseline", "_ref", "CybORG", "Evaluation")
    if not os.path.exists(os.path.join(ref_eval, "evaluation.py")):
        pytest.skip("baseline/_ref is not installed on this box")
    code = ("import importlib.util, sys\n"
            f"spec = importlib.util.spec_from_file_location('ref_evaluation', {os.path.join(ref_eval, 'evaluation.py')!r})\n"
            "m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)\n"
            "import CybORG; assert 'compat' in CybORG.__file__, CybORG.__file__\n"
            "sub = m.load_submission(sys.argv[2])\n"
            "m.run_evaluation(sub, log_path=sys.argv[1], max_eps=2, write_to_file=True, seed=7)\n")
    # a submission in the layout of the reference's example_submission/ (submission.py + dummy_agent.py, same classes and imports; the
    # example itself cannot be imported by any Python: its `from __future__ import annotations` sits on line 9, submission.py:9)
    submission_py = ("from __future__ import annotations\n"
                     "from CybORG import CybORG\nfrom CybORG.Agents import BaseAgent\n"
                     "from ray.rllib.env.multi_agent_env import MultiAgentEnv\nfrom CybORG.Agents.Wrappers.EnterpriseMAE import EnterpriseMAE\n"
                     "from dummy_agent import DummyAgent\n\n"
                     "class Submission:\n    NAME: str = 'SUBMISSION NAME'\n    TEAM: str = 'TEAM NAME'\n    TECHNIQUE: str = 'TECHNIQUE NAME'\n"
                     "    AGENTS: dict[str, BaseAgent] = {f'blue_agent_{agent}': DummyAgent() for agent in range(5)}\n\n"
                     "    def wrap(env: CybORG) -> MultiAgentEnv:\n        return EnterpriseMAE(env)\n")
    dummy_py = open(os.path.join(ref_eval, "example_submission", "dummy_agent.py")).read()
    # the example submission imports ray's MultiAgentEnv and gym's Space for type annotations: stand-ins from the test infrastructure
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(root, "compat"), os.path.join(root, "oracle", "refshim")]))
    with tempfile.TemporaryDirectory() as tmp, tempfile.TemporaryDirectory() as subdir:
        open(os.path.join(subdir, "submission.py"), "w").write(submission_py)
        open(os.path.join(subdir, "dummy_agent.py"), "w").write(dummy_py)
        out = subprocess.run([sys.executable, "-c", code, tmp + "/", subdir], capture_output=True, text=True, env=env, cwd=tmp)
        assert out.returncode == 0, out.stderr[-1500:]
        assert "Average reward is:" in out.stdout
        assert sorted(os.listdir(tmp)) == ["actions.txt", "full.txt", "scores.txt", "summary.json", "summary.txt"]
        summary = json.load(open(os.path.join(tmp, "summary.json")))
        assert summary["parameters"] == {"seed": 7, "episode_length": 500, "max_episodes": 2} and summary["reward"]["mean"] < 0
