"""Canonical flattened CC4 episode state (TEST INFRASTRUCTURE — never imported by the product package).

One int32 stream per (episode, tick) that the reference harness (oracle/ref_harness.py), the C oracle
(oracle/cc4_oracle.c: cc4o_export) and the CUDA engine (cc4_export_canon in include/cc4_b200.h) all
produce in exactly this layout, so that "state parity" is a plain array comparison.  `decode` turns a
stream back into a dict for readable diffs.

Stream layout (all int32):
  phase, step, blocks[9]                       blocks[to] bit from = state.blocks[to] contains from   (State.py:148)
  n_hosts, then per existing host in slot order:
      slot, evbits(1 proc_creation | 2 net_conn | 4 old_proc | 8 old_net), nproc, (pid, kind, root)*nproc,
      nsvc, (code, pid, active, reliability)*nsvc                               (Host.py:117-124, Service.py:15)
  per red agent 0..5:
      active, sessions_count, nsess, (ident, slot, is_red_abstract, is_root, pid)*nsess   in dict order (State.py:305)
      n, ip-known slots sorted; n, hostname-known slots sorted; subnet-known mask        (ActionSpace.py:170)
      n, server_session ids that are True, in dict order
      n, (slot, fsm_state, hostname_known)*n   FiniteStateRedAgent.host_states in insertion order (slot -1 = None key),
                                                as it stands once the agent has digested this tick's observation
                                                (FiniteStateRedAgent.py:91-93; the reference does that lazily at the
                                                start of the next tick, the engines do it eagerly)
      n, (ident, slot, portmask)*n sorted      RedAbstractSession.ports                   (Session.py:125)
      inprog_code, inprog_remaining            (-1, 0 when nothing in flight)             (SimulationController.py:246)
  per blue agent 0..4:
      n, (slot, pid)*n  VelociraptorServer.sus_pids, stable-sorted by slot               (Session.py:160)
      inprog_code, inprog_remaining

Slots: regular hosts subnet*16 + (k for server_host_k | 6+k for user_host_k) with subnets in creation
order (EnterpriseScenarioGenerator.py:40-51); routers 128+subnet; root_internet_host_0 = 136.
"""
from __future__ import annotations

SUBNETS_CREATION = [
    "restricted_zone_a_subnet", "operational_zone_a_subnet", "restricted_zone_b_subnet",
    "operational_zone_b_subnet", "contractor_network_subnet", "public_access_zone_subnet",
    "admin_network_subnet", "office_network_subnet", "internet_subnet",
]
SUBNETS_SORTED = sorted(SUBNETS_CREATION)
SUBNET_IDX = {n: i for i, n in enumerate(SUBNETS_CREATION)}
N_SLOTS = 137
SLOT_ROUTER_BASE = 128
SLOT_INTERNET = 136

# process kinds
PK = dict(SSHD=0, OT=1, APACHE2=2, MYSQLD=3, SMTP=4, D_APACHE=5, D_TOMCAT=6, D_HARAKA=7, D_VSFTPD=8,
          SESSION=9, SHELL=10, REDABS=11)
SERVICE_NAMES = ["SSHD", "OTSERVICE", "APACHE2", "MYSQLD", "SMTP", "apache2", "tomcat", "haraka", "vsftpd"]
FSM_STATES = ["K", "KD", "S", "SD", "U", "UD", "R", "RD", "F"]
PORT_BIT = {22: 0, 1: 1, 80: 2, 3390: 3, 25: 4, 443: 5}

ACTION_CODE = {
    "Sleep": 0, "Monitor": 1, "Analyse": 2, "Remove": 3, "Restore": 4, "DeployDecoy": 5,
    "BlockTrafficZone": 6, "AllowTrafficZone": 7, "InvalidAction": 8,
    "DiscoverRemoteSystems": 10, "AggressiveServiceDiscovery": 11, "StealthServiceDiscovery": 12,
    "DiscoverDeception": 13, "ExploitRemoteService": 14, "PrivilegeEscalate": 15, "Impact": 16,
    "DegradeServices": 17, "Withdraw": 18, "GreenAccessService": 20, "GreenLocalWork": 21,
}


def host_slot(hostname: str) -> int:
    if hostname == "root_internet_host_0":
        return SLOT_INTERNET
    for sn, si in SUBNET_IDX.items():
        if hostname.startswith(sn + "_"):
            rest = hostname[len(sn) + 1:]
            if rest == "router":
                return SLOT_ROUTER_BASE + si
            kind, _, k = rest.rpartition("_")
            k = int(k)
            if kind == "server_host":
                return si * 16 + k
            if kind == "user_host":
                return si * 16 + 6 + k
    raise ValueError(hostname)


def slot_name(slot: int) -> str:
    if slot == SLOT_INTERNET:
        return "root_internet_host_0"
    if slot >= SLOT_ROUTER_BASE:
        return SUBNETS_CREATION[slot - SLOT_ROUTER_BASE] + "_router"
    si, k = divmod(slot, 16)
    return SUBNETS_CREATION[si] + (f"_server_host_{k}" if k < 6 else f"_user_host_{k - 6}")


def encode(st: dict) -> list[int]:
    out: list[int] = [st["phase"], st["step"]] + list(st["blocks"])
    hosts = st["hosts"]
    out.append(len(hosts))
    for slot in sorted(hosts):
        h = hosts[slot]
        ev = h["ev"]
        out += [slot, ev[0] | ev[1] << 1 | ev[2] << 2 | ev[3] << 3, len(h["procs"])]
        for p in h["procs"]:
            out += list(p)
        out.append(len(h["svcs"]))
        for s in h["svcs"]:
            out += list(s)
    for r in st["red"]:
        out += [r["active"], r["count"], len(r["sessions"])]
        for s in r["sessions"]:
            out += list(s)
        out += [len(r["ip_known"])] + list(r["ip_known"])
        out += [len(r["host_known"])] + list(r["host_known"])
        out.append(r["subnet_known"])
        out += [len(r["server_sessions"])] + list(r["server_sessions"])
        out.append(len(r["fsm"]))
        for f in r["fsm"]:
            out += list(f)
        out.append(len(r["ports"]))
        for p in r["ports"]:
            out += list(p)
        out += list(r["inprog"])
    for b in st["blue"]:
        out.append(len(b["sus"]))
        for s in b["sus"]:
            out += list(s)
        out += list(b["inprog"])
    return [int(x) for x in out]


class _Rd:
    def __init__(self, a):
        self.a, self.i = a, 0

    def one(self):
        v = int(self.a[self.i])
        self.i += 1
        return v

    def tup(self, n):
        v = tuple(int(x) for x in self.a[self.i:self.i + n])
        self.i += n
        return v


def decode(stream) -> dict:
    r = _Rd(stream)
    st: dict = {"phase": r.one(), "step": r.one(), "blocks": list(r.tup(9))}
    hosts = {}
    for _ in range(r.one()):
        slot, ev, nproc = r.tup(3)
        procs = [r.tup(3) for _ in range(nproc)]
        svcs = [r.tup(4) for _ in range(r.one())]
        hosts[slot] = dict(ev=(ev & 1, ev >> 1 & 1, ev >> 2 & 1, ev >> 3 & 1), procs=procs, svcs=svcs)
    st["hosts"] = hosts
    red = []
    for _ in range(6):
        active, count, ns = r.tup(3)
        sessions = [r.tup(5) for _ in range(ns)]
        ipk = list(r.tup(r.one()))
        hnk = list(r.tup(r.one()))
        snk = r.one()
        ssk = list(r.tup(r.one()))
        fsm = [r.tup(3) for _ in range(r.one())]
        ports = [r.tup(3) for _ in range(r.one())]
        inprog = r.tup(2)
        red.append(dict(active=active, count=count, sessions=sessions, ip_known=ipk, host_known=hnk,
                        subnet_known=snk, server_sessions=ssk, fsm=fsm, ports=ports, inprog=inprog))
    st["red"] = red
    blue = []
    for _ in range(5):
        sus = [r.tup(2) for _ in range(r.one())]
        blue.append(dict(sus=sus, inprog=r.tup(2)))
    st["blue"] = blue
    assert r.i == len(stream), (r.i, len(stream))
    return st


def diff(a: dict, b: dict, path="") -> list[str]:
    """Readable differences between two decoded states (a = expected, b = got)."""
    out = []
    if isinstance(a, dict) and isinstance(b, dict):
        for k in sorted(set(a) | set(b), key=str):
            if k not in a:
                out.append(f"{path}/{k}: unexpected {b[k]}")
            elif k not in b:
                out.append(f"{path}/{k}: missing {a[k]}")
            else:
                out += diff(a[k], b[k], f"{path}/{k}")
    elif isinstance(a, list) and isinstance(b, list) and len(a) == len(b) and a and isinstance(a[0], dict):
        for i, (x, y) in enumerate(zip(a, b)):
            out += diff(x, y, f"{path}[{i}]")
    elif a != b:
        out.append(f"{path}: expected {a} got {b}")
    return out
