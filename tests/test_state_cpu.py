"""state.true_state (product-side decoder of cc4_export_canon's stream) against the test infrastructure's canon.decode, on
streams of the C oracle (no GPU needed)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def test_true_state_matches_canon_decode():
    import canon
    import oracle_py
    from cage_challenge_4_b200.state import hostname, true_state, true_state_table, PROCESS_KINDS
    env = oracle_py.OracleEnv(seed=5, episode=2, steps=500)
    env.reset()
    g = np.random.default_rng(0)
    for t in range(150):
        acts = np.array([g.integers(0, k) for k in (82, 82, 82, 82, 242)], dtype=np.int32)
        env.step(acts, np.zeros(5, dtype=np.uint8))
    stream = env.export()
    want, got = canon.decode(stream), true_state(stream)
    assert got["mission_phase"] == want["phase"] and got["step"] == want["step"] == 150
    assert set(got["hosts"]) == {canon.slot_name(s) for s in want["hosts"]}
    for slot, h in want["hosts"].items():
        mine = got["hosts"][hostname(slot)]
        assert hostname(slot) == canon.slot_name(slot)
        assert [(p["PID"], PROCESS_KINDS.index(p["process_name"]), int(p["username"] == "root")) for p in mine["Processes"]] == [tuple(p) for p in h["procs"]]
        assert [(PROCESS_KINDS.index(s["service_name"]), s["PID"], int(s["active"]), s["reliability"]) for s in mine["Services"]] == [tuple(s) for s in h["svcs"]]
        assert tuple(int(mine["events"][k]) for k in ("process_creation", "network_connections", "old_process_creation", "old_network_connections")) == h["ev"]
    n_sessions = 0
    for r, rr in enumerate(want["red"]):
        mine = got["red"][f"red_agent_{r}"]
        assert [(s["session_id"], s["hostname"]) for s in mine["sessions"]] == [(s[0], canon.slot_name(s[1])) for s in rr["sessions"]]
        assert [v["state"] for v in mine["host_states"].values()] == [canon.FSM_STATES[f[1]] for f in rr["fsm"]]
        n_sessions += len(rr["sessions"])
    assert n_sessions > 3 and sum(len(h["Sessions"]) for h in got["hosts"].values()) == n_sessions
    rows = true_state_table(got)
    assert len(rows) == len(want["hosts"]) and any(r["Access"] != "None" for r in rows)


class _OracleBackedFacade:
    """Stands in for wrappers.CybORG in CPU tests: the same attributes the observation tracker reads, filled from the C oracle
    in production (Philox) mode instead of the CUDA engine."""

    def __init__(self, seed, episode, steps):
        import numpy as np
        import torch
        import oracle_py as O
        self.o = O.OracleEnv(seed=seed, episode=episode, steps=steps)
        self.o.reset()
        self.engine = self
        self._mask = torch.from_numpy(self.o.mask().astype(bool))[None]
        self._obs = torch.from_numpy(self.o.observe())[None]
        self._last_actions = None
        self.np = np

    def export_canon(self, k):
        return self.o.export()

    def tick(self, acts, msgs):
        import torch
        rc, obs, r, d = self.o.step(self.np.asarray(acts, dtype=self.np.int32), self.np.asarray(msgs, dtype=self.np.uint8))
        self._last_actions = self.np.asarray(acts, dtype=self.np.int32)[None]
        prev = self._obs[0].numpy()
        self._obs = torch.from_numpy(obs)[None]
        return prev


def _check_dict_observations(name, make_facade):
    import json
    import numpy as np
    from cage_challenge_4_b200.observations import BlueObservationTracker, blue_observation
    from cage_challenge_4_b200.wrappers import BlueFixedActionWrapper
    fx = np.load(os.path.join(ROOT, "tests", "golden", name))
    expected = json.loads(str(fx["summaries"]))
    cy, tick = make_facade(int(fx["seed"]), int(fx["episode"]), int(fx["steps"]))
    helper = BlueFixedActionWrapper.__new__(BlueFixedActionWrapper)
    helper.env, helper._pad_spaces, helper.env_index = cy, False, 0
    helper._refresh_masks()
    assert np.array_equal(cy._mask[0].numpy().astype(np.uint8), fx["mask"])
    cy._tracker = BlueObservationTracker(cy)
    for t in range(len(fx["actions"])):
        prev = tick(fx["actions"][t], fx["msgs"][t])
        cy._tracker.after_tick(prev, helper)
        for b in range(5):
            got, exp = blue_observation(cy, f"blue_agent_{b}"), expected[t][b]
            where = f"{name} tick {t} blue_agent_{b}"
            assert got["success"].name == exp["success"], f"{where}: success {got['success'].name} != {exp['success']}"
            if exp["success"] != "IN_PROGRESS":
                assert str(got["action"]) == exp["action"], f"{where}: action {got['action']} != {exp['action']}"
            hosts = {k: sorted(p["PID"] for p in v.get("Processes", []) if "PID" in p) for k, v in got.items() if k not in ("success", "action", "message")}
            assert hosts == exp["hosts"], f"{where}: hosts {hosts} != {exp['hosts']}"
            sent = [m for j, m in enumerate(fx["msgs"][t]) if j != b]
            assert [int(sum(int(x) << i for i, x in enumerate(m))) for m in got["message"]] == [int(x) for x in sent], where


def test_dict_observations_match_the_reference_cpu():
    """observations.py (dict observations rebuilt from the canonical state) against the LIVE reference's get_observation, on
    the Philox seam: tests/golden/dictobs_*.npz hold, per tick and blue agent, success / action / reported hosts / pids of
    the reference (oracle/ref_harness.py --dictobs).  Here the state stream comes from the C oracle; the GPU twin of this
    test (tests/test_wrappers_gpu.py) drives the real facade."""
    for name in sorted(f for f in os.listdir(os.path.join(ROOT, "tests", "golden")) if f.startswith("dictobs_")):
        def make(seed, episode, steps):
            f = _OracleBackedFacade(seed, episode, steps)
            return f, f.tick
        _check_dict_observations(name, make)
