"""The mirrored wrapper API, checked the way the reference checks its own wrappers
(CybORG/Tests/test_cc4/test_BlueEnterpriseWrapper.py, test_mission_phase.py, test_cc4_seed.py)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from cage_challenge_4_b200.wrappers import (BlueEnterpriseWrapper, BlueFlatWrapper, CybORG, EnterpriseGreenAgent,  # noqa: E402
                                            EnterpriseMAE, EnterpriseScenarioGenerator, FiniteStateRedAgent, SleepAgent,
                                            SUBNETS_SORTED)

NUM_SUBNETS, MAX_HOSTS, MESSAGE_LENGTH, NUM_MESSAGES = 9, 16, 8, 4


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
    blk = labels.index("BlockTrafficZone restricted_zone_a_subnet <- contractor_network_subnet")
    alw = labels.index("AllowTrafficZone restricted_zone_a_subnet <- contractor_network_subnet")
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


def test_pad_spaces():
    env = make(cls=BlueFlatWrapper, pad_spaces=True)
    obs, info = env.reset()
    assert all(len(v) == 210 for v in obs.values()) and all(len(i["action_mask"]) == 242 for i in info.values())
    obs, rew, *_ = env.step({"blue_agent_0": 241})       # padding index acts as Sleep
    assert all(len(v) == 210 for v in obs.values())


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

    out = run_evaluation(Submission, max_eps=2, seed=3)
    assert out["episodes"] == 2 and out["mean"] < 0
    pol = MLPPolicy().cuda()
    b1 = run_evaluation_batched(pol, max_eps=16, seed=5)
    lo = run_evaluation_batched(pol, max_eps=16, seed=5, rank=0, world=1)
    assert b1["rewards"] == lo["rewards"] and b1["episodes"] == 16 and b1["stdev"] > 0


def test_random_blue_policy_picks_valid_non_traffic_actions():
    """cc4BlueRandomAgent equivalent (RandomAgent.py:66-69) on device: valid actions only, no Block / Allow."""
    import torch
    from cage_challenge_4_b200.engine import BatchedCC4
    from cage_challenge_4_b200.evaluation import RandomBluePolicy, run_evaluation_batched
    env = BatchedCC4(256, device=0, seed=5)
    obs, mask = env.reset()
    pol = RandomBluePolicy(seed=1)
    seen = set()
    for _ in range(20):
        a = pol(obs, mask)
        assert bool(torch.gather(mask, 2, a.long().unsqueeze(-1)).all())
        assert not bool(((a[:, :4] >= 50) & (a[:, :4] < 66)).any()) and not bool(((a[:, 4] >= 146) & (a[:, 4] < 194)).any())
        seen |= set(a[:, 0].tolist())
        obs, rew, done = env.step(a)
    assert len(seen) > 20
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
    with pytest.raises(NotImplementedError):
        env.env.get_agent_state("blue_agent_0")
