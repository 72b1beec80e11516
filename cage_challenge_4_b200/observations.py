"""Per-agent dict observations of ONE episode, rebuilt from the engine's state.

The reference hands every agent a nested dict (`CybORG.get_observation`, env.py:245; built by Shared/Observation.py from the
Observation objects of the tick's actions).  The engine keeps only the FACTS that the scripted agents, the reward and the
blue vectors consume, so the dict is reconstructed on the host, for one selected episode, from the canonical state export
(`cc4_export_canon`) and what the facade already has: the indices submitted this tick, the previous observation vector
and the masks.  For a blue agent the dict holds

  success   TernaryEnum of the agent's own action this tick (Enums.py:5): IN_PROGRESS while a multi-tick action runs
            (SimulationController.py:256-265), UNKNOWN for Sleep (Action.py:52), FALSE for Block / Allow that change nothing
            (ControlTraffic.py:62-116, 132-185), TRUE otherwise (a blue agent always holds sessions on the hosts its indices
            can name, so Analyse / Remove / Restore / DeployDecoy / Monitor succeed: Analyse.py:40, Remove.py:42, ...)
  action    the action that came due this tick as it prints in the reference (`Restore <hostname>`), Sleep while waiting
  message   the four other agents' 8-bit messages in sender order (SimulationController.send_messages :781-804)
  <host>    one entry per host of the agent's subnets (routers included) on which this tick's end-turn Monitor found events
            (Monitor.py:35-74): `Interface` [{ip_address}], `System info` {Hostname, OSType, ...}, and `Processes`: one
            {'PID': pid} per process-creation event that carries a pid (these are the pids Remove will act on,
            Session.py:154-165), {'local_address': ip} for one that does not (GreenLocalWork false positive,
            GreenLocalWork.py:112-116), one {'Connections': [{'local_address': ip}]} when there were network-connection events
            (the engine keeps one bit per host for those: ports, remote ends and the number of events are not tracked);
            plus, when the agent's DeployDecoy completed this tick, the decoy process on its host (DecoyAction.py:86-112).
            and, when its Analyse completed on a host where red has left files, their listing (`Files`: cmd.sh per exploit,
            escalate.sh per privilege escalation since the host's last Restore; the tracker derives them from the shells and
            root flags it sees appear, the engine itself keeps no files).
Not reconstructed: operating-system details (fixed labels), remote addresses and ports of connections, file metadata beyond
name / density / signature.
tests/test_wrappers_gpu.py::test_dict_observations_against_the_reference compares success, action, the reported hosts and the
process-creation pids with the live reference on a golden tape.
"""
from __future__ import annotations

import numpy as np

from .state import SUBNETS, hostname as slot_hostname


def host_slot(name: str) -> int:
    """Inverse of state.hostname."""
    if name == "root_internet_host_0":
        return 136
    for i, sn in enumerate(SUBNETS):
        if name.startswith(sn + "_"):
            rest = name[len(sn) + 1:]
            if rest == "router":
                return 128 + i
            kind, _, k = rest.rpartition("_")
            return i * 16 + (int(k) if kind == "server_host" else 6 + int(k))
    raise KeyError(name)

# action codes of the canonical export (oracle/canon.py ACTION_CODE)
_CODE_NAME = {0: "Sleep", 1: "Monitor", 2: "Analyse", 3: "Remove", 4: "Restore", 5: "DeployDecoy", 6: "BlockTrafficZone", 7: "AllowTrafficZone"}
_DECOY_SERVICE = {5: "apache2", 6: "tomcat", 7: "haraka", 8: "vsftpd"}
_DURATION = {"Analyse": 2, "Remove": 3, "Restore": 5, "DeployDecoy": 2}      # Analyse.py:38, Remove.py:40, Restore.py:36, DeployDecoy.py:31; the rest 1


def _decode(stream):
    """The parts of the canonical stream this module needs (layout: state.py): blocks, per host event bits and processes,
    per blue agent suspicious pids and the in-flight action."""
    a, i = [int(x) for x in stream], 2
    blocks = a[i:i + 9]
    i += 9
    hosts = {}
    nh = a[i]
    i += 1
    for _ in range(nh):
        slot, ev, nproc = a[i:i + 3]
        i += 3
        procs = [tuple(a[i + 3 * k:i + 3 * k + 3]) for k in range(nproc)]
        i += 3 * nproc
        nsvc = a[i]
        i += 1 + 4 * nsvc
        hosts[slot] = (ev, procs)
    for _ in range(6):
        i += 2
        ns = a[i]
        i += 1 + 5 * ns
        for _ in range(2):
            i += 1 + a[i]
        i += 1
        i += 1 + a[i]
        i += 1 + 3 * a[i]
        i += 1 + 3 * a[i]
        i += 2
    blue = []
    for _ in range(5):
        n = a[i]
        sus = [(a[i + 1 + 2 * k], a[i + 2 + 2 * k]) for k in range(n)]
        i += 1 + 2 * n
        blue.append((sus, a[i], a[i + 1]))
        i += 2
    assert i == len(a)
    return blocks, hosts, blue


class BlueObservationTracker:
    """Follows one episode tick by tick (the facade calls `after_reset` / `after_tick`): which action of every blue agent is in
    flight (the duration queue, SimulationController.py:245-265) and which suspicious pids are new."""

    def __init__(self, cyborg, env_index: int = 0):
        self.cy, self.k = cyborg, env_index
        self.after_reset()

    def _export(self):
        return _decode(self.cy.engine.export_canon(self.k))

    def after_reset(self):
        self._per_slot = [dict() for _ in range(5)]     # suspicious pids seen so far, per agent and host slot (empty at reset)
        # Files red leaves behind (what Analyse lists): an exploit drops cmd.sh next to the shell it opens (ExploitAction.py:228-238),
        # a privilege escalation drops escalate.sh (EscalateAction.py:70-77); only Restore removes them (Host.restore Host.py:373-380).
        # The engine does not keep files; the tracker derives them from the shells and root flags it sees appear.
        self.files: dict[int, list[str]] = {}
        self._red_procs = {slot: {pid: root for pid, kind, root in procs if kind in (10, 11)} for slot, (ev, procs) in self._export()[1].items()}
        self.inflight = [None] * 5          # (descriptor, ticks left)
        self.current = [None] * 5           # what get_observation reports for the last tick

    def after_tick(self, prev_obs_row: np.ndarray, helper):
        """prev_obs_row: the episode's observation vectors BEFORE the tick (blocked bits for Block / Allow success)."""
        from .wrappers import ActionDescriptor, BLUE_SUBNETS, SUBNETS_CREATION, SUBNETS_SORTED, TernaryEnum, host_ip
        blocks, hosts, blue = self._export()
        acts = self.cy._last_actions[0]
        out = []
        files_before = {k: list(v) for k, v in self.files.items()}      # blue acts before red: Analyse sees last tick's files
        restored = []
        for b in range(5):
            agent = f"blue_agent_{b}"
            if self.inflight[b] is None:
                d = helper.actions(agent)[int(acts[b])]
                self.inflight[b] = [d, _DURATION.get(d.name, 1)]
            d, left = self.inflight[b]
            left -= 1
            obs: dict = {}
            if left >= 1:
                self.inflight[b][1] = left
                success, action = TernaryEnum.IN_PROGRESS, ActionDescriptor("Sleep")
            else:
                self.inflight[b] = None
                action = d
                if d.name == "Sleep":
                    success = TernaryEnum.UNKNOWN
                elif d.name in ("BlockTrafficZone", "AllowTrafficZone"):
                    own = sorted(BLUE_SUBNETS[b]).index(d.params["to_subnet"])
                    was_blocked = bool(prev_obs_row[b, 1 + 59 * own + 9 + SUBNETS_SORTED.index(d.params["from_subnet"])])
                    success = TernaryEnum.TRUE if was_blocked == (d.name == "AllowTrafficZone") else TernaryEnum.FALSE
                else:
                    success = TernaryEnum.TRUE
                if d.name == "Restore":
                    restored.append(host_slot(d.params["hostname"]))
                if d.name == "Analyse" and files_before.get(host_slot(d.params["hostname"])):     # DensityScout + SigCheck (Analyse.py:56-71)
                    obs[d.params["hostname"]] = {"Files": [{"File Name": f, "Path": "/tmp/", "Density": 0.9, "Signed": False}
                                                           for f in files_before[host_slot(d.params["hostname"])]]}
                if d.name == "DeployDecoy":          # the new decoy process: the last decoy in the host's list (DecoyAction.py:86-112)
                    slot = host_slot(d.params["hostname"])
                    decoys = [(pid, kind) for pid, kind, _ in hosts.get(slot, (0, []))[1] if kind in _DECOY_SERVICE]
                    if decoys:
                        pid, kind = decoys[-1]
                        obs[d.params["hostname"]] = {"Processes": [{"PID": pid, "PPID": 1, "service_name": _DECOY_SERVICE[kind], "username": "ubuntu",
                                                                    "Properties": ["rfi"]}]}
            sus = blue[b][0]
            fresh = self._fresh(b, sus)              # suspicious pids appended this tick, by host slot
            monitored = []
            for sn in BLUE_SUBNETS[b]:
                si = SUBNETS_CREATION.index(sn)
                monitored += [128 + si] + [si * 16 + k for k in range(16)]
            for slot in monitored:
                if slot not in hosts:
                    continue
                ev = hosts[slot][0]
                if not ev & 12:          # old_process_creation | old_network_connections: what this tick's Monitor reported
                    continue
                h = slot_hostname(slot)
                ip = host_ip(h)
                procs = []
                if ev & 8:
                    procs.append({"Connections": [{"local_address": ip}]})
                if ev & 4:
                    pids = fresh.get(slot, [])
                    procs += [{"PID": p} for p in pids] or [{"local_address": ip}]
                entry = obs.setdefault(h, {})
                entry["Interface"] = [{"ip_address": ip}]
                entry["Processes"] = entry.get("Processes", []) + procs
                entry["System info"] = {"Hostname": h, "OSType": "LINUX", "OSDistribution": "UBUNTU", "Architecture": "x64"}
            obs["success"], obs["action"] = success, action
            out.append(obs)
        self.current = out
        # files: this tick's Restores first (blue phase), then what the red phase dropped
        for slot in restored:
            self.files.pop(slot, None)
            self._red_procs[slot] = {}
        for slot, (ev, procs) in hosts.items():
            now = {pid: root for pid, kind, root in procs if kind in (10, 11)}
            was = self._red_procs.get(slot, {})
            for pid, root in now.items():
                if pid not in was and any(k == 10 and p == pid for p, k, _ in procs):
                    self.files.setdefault(slot, []).append("cmd.sh")
                if root and not was.get(pid, 0):
                    self.files.setdefault(slot, []).append("escalate.sh")
            self._red_procs[slot] = now

    def _fresh(self, b, sus):
        """{slot: [pid, ...]} of the entries appended since the last tick.  The export lists an agent's entries grouped by
        slot (stable), so "appended" is recovered per slot from the counts seen last tick."""
        prev = self._per_slot
        cur: dict = {}
        for slot, pid in sus:
            cur.setdefault(slot, []).append(pid)
        fresh = {slot: pids[len(prev[b].get(slot, [])):] for slot, pids in cur.items() if len(pids) > len(prev[b].get(slot, []))}
        self._per_slot[b] = cur
        return fresh


def blue_observation(cyborg, agent: str, env_index: int = 0) -> dict:
    """The dict `CybORG.get_observation(agent)` returns for a blue agent (see the module docstring)."""
    tr = cyborg._tracker
    b = int(agent.rsplit("_", 1)[1])
    from .wrappers import TernaryEnum
    msgs = cyborg._obs[env_index, b].cpu().numpy()
    n = 210 if b == 4 else 92
    message = [msgs[n - 32 + 8 * k:n - 24 + 8 * k].astype(bool) for k in range(4)]
    if tr is None or tr.current[b] is None:       # right after reset: the initial observation (SimulationController.py:182-196)
        return {"success": TernaryEnum.UNKNOWN, "message": message}
    obs = dict(tr.current[b])
    obs["message"] = message
    return obs
