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
