"""True-state view of one episode for debugging and logs (the role of `CybORG.get_agent_state('True')`, env.py:265, and of
`TrueStateTableWrapper`, Agents/Wrappers/TrueStateWrapper.py): decodes the int32 stream `cc4_export_canon` writes.

Stream layout (include/cc4_b200.h: cc4_export_canon):
  phase, step, blocks[9]                      blocks[to] bit from = traffic from -> to is blocked   (State.py:148)
  n_hosts, per existing host in slot order:   slot, event bits (1 process_creation | 2 network_connections | 4 / 8 the
                                              same of the previous tick), nproc, (pid, kind, root) * nproc,
                                              nsvc, (code, pid, active, reliability) * nsvc
  per red agent 0..5:                         active, sessions_count, nsess, (ident, slot, abstract, root, pid) * nsess,
                                              n + known ip slots, n + known hostname slots, known-subnet mask,
                                              n + server-session idents, n + (slot, fsm state, hostname known) * n,
                                              n + (ident, slot, port mask) * n, in-flight (action code, ticks left)
  per blue agent 0..4:                        n + (slot, pid) * n suspicious pids, in-flight (action code, ticks left)
Host slots: subnet * 16 + (k for server_host_k | 6 + k for user_host_k), subnets in creation order
(EnterpriseScenarioGenerator.py:40-51); routers 128 + subnet; root_internet_host_0 = 136.
"""
from __future__ import annotations

SUBNETS = ["restricted_zone_a_subnet", "operational_zone_a_subnet", "restricted_zone_b_subnet", "operational_zone_b_subnet",
           "contractor_network_subnet", "public_access_zone_subnet", "admin_network_subnet", "office_network_subnet", "internet_subnet"]
PROCESS_KINDS = ["sshd", "otservice", "apache2", "mysqld", "smtp", "decoy apache2", "decoy tomcat", "decoy haraka", "decoy vsftpd",
                 "session (blue / green)", "red shell", "red abstract session"]
FSM_STATES = ["K", "KD", "S", "SD", "U", "UD", "R", "RD", "F"]
ACTIONS = {0: "Sleep", 1: "Monitor", 2: "Analyse", 3: "Remove", 4: "Restore", 5: "DeployDecoy", 6: "BlockTrafficZone", 7: "AllowTrafficZone",
           8: "InvalidAction", 10: "DiscoverRemoteSystems", 11: "AggressiveServiceDiscovery", 12: "StealthServiceDiscovery",
           13: "DiscoverDeception", 14: "ExploitRemoteService", 15: "PrivilegeEscalate", 16: "Impact", 17: "DegradeServices", 18: "Withdraw"}
PORTS = [22, 1, 80, 3390, 25, 443]          # bit i of a port mask


def hostname(slot: int) -> str:
    if slot == 136:
        return "root_internet_host_0"
    if slot >= 128:
        return SUBNETS[slot - 128] + "_router"
    sn, k = divmod(slot, 16)
    return SUBNETS[sn] + (f"_server_host_{k}" if k < 6 else f"_user_host_{k - 6}")


def true_state(stream) -> dict:
    """Nested dict in the spirit of the reference's true-state observation: hostnames -> Processes / Services / Sessions /
    events, plus the red agents' knowledge and FSM states, the blue agents' suspicious pids, blocks and mission phase."""
    a, i = [int(x) for x in stream], 0

    def take(n):
        nonlocal i
        v = a[i:i + n]
        i += n
        return v

    phase, step = take(2)
    blocks = take(9)
    out: dict = {"mission_phase": phase, "step": step,
                 "blocked": {SUBNETS[to]: [SUBNETS[f] for f in range(9) if blocks[to] >> f & 1] for to in range(9) if blocks[to]},
                 "hosts": {}, "red": {}, "blue": {}}
    for _ in range(take(1)[0]):
        slot, ev, nproc = take(3)
        procs = []
        for _ in range(nproc):
            pid, kind, root = take(3)
            procs.append({"PID": pid, "process_name": PROCESS_KINDS[kind], "username": "root" if root else "user"})
        svcs = []
        for _ in range(take(1)[0]):
            code, pid, active, rel = take(4)
            svcs.append({"service_name": PROCESS_KINDS[code], "PID": pid, "active": bool(active), "reliability": rel})
        out["hosts"][hostname(slot)] = {"Processes": procs, "Services": svcs, "Sessions": [],
                                        "events": {"process_creation": bool(ev & 1), "network_connections": bool(ev & 2),
                                                   "old_process_creation": bool(ev & 4), "old_network_connections": bool(ev & 8)}}
    for r in range(6):
        active, count, ns = take(3)
        name = f"red_agent_{r}"
        sessions = []
        for _ in range(ns):
            ident, slot, abstract, root, pid = take(5)
            s = {"agent": name, "session_id": ident, "hostname": hostname(slot), "username": "root" if root else "user", "PID": pid,
                 "Type": "RedAbstractSession" if abstract else "shell"}
            sessions.append(s)
            out["hosts"][hostname(slot)]["Sessions"].append(s)
        ipk = [hostname(s) for s in take(take(1)[0])]
        hnk = [hostname(s) for s in take(take(1)[0])]
        snk = take(1)[0]
        ssk = take(take(1)[0])
        fsm = {}
        for _ in range(take(1)[0]):
            slot, st, hn = take(3)
            fsm[hostname(slot) if slot >= 0 else None] = {"state": FSM_STATES[st], "hostname_known": bool(hn)}
        ports = []
        for _ in range(take(1)[0]):
            ident, slot, mask = take(3)
            ports.append({"session_id": ident, "hostname": hostname(slot), "ports": [p for b, p in enumerate(PORTS) if mask >> b & 1]})
        code, left = take(2)
        out["red"][name] = {"active": bool(active), "sessions_count": count, "sessions": sessions, "known_ips": ipk, "known_hostnames": hnk,
                            "known_subnets": [SUBNETS[s] for s in range(9) if snk >> s & 1], "server_sessions": ssk, "host_states": fsm,
                            "known_ports": ports, "in_progress": None if code < 0 else {"action": ACTIONS.get(code, str(code)), "ticks_left": left}}
    for b in range(5):
        sus = [{"hostname": hostname(s), "PID": p} for s, p in (take(2) for _ in range(take(1)[0]))]
        code, left = take(2)
        out["blue"][f"blue_agent_{b}"] = {"suspicious_pids": sus,
                                          "in_progress": None if code < 0 else {"action": ACTIONS.get(code, str(code)), "ticks_left": left}}
    assert i == len(a), "state stream has trailing values"
    return out


def true_state_table(state: dict) -> list[dict]:
    """One row per host, like TrueStateTableWrapper's table (TrueStateWrapper.py): hostname, processes, services up, red sessions."""
    rows = []
    for name, h in state["hosts"].items():
        red = h["Sessions"]
        rows.append({"Hostname": name, "Processes": len(h["Processes"]),
                     "Services": ", ".join(s["service_name"] for s in h["Services"] if s["active"]),
                     "Access": "Privileged" if any(s["username"] == "root" for s in red) else "User" if red else "None",
                     "Red sessions": ", ".join(f'{s["agent"]}:{s["session_id"]}' for s in red)})
    return rows
