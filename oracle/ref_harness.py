"""Reference-side harness (TEST INFRASTRUCTURE — never imported by the product package).

Runs the UNMODIFIED Python reference from /root/reference (with the import stand-ins in oracle/refshim)
and produces, for one episode:
  * an RNG *tape*: every draw the reference makes, tagged with its call site (file:line) and the acting
    agent ("actor"), in the order the reference made them;
  * a canonical flattened state after reset and after every step (the comparison key for oracle + CUDA);
  * the 5 blue observation vectors, the team reward and done flag of every step.

The reference accepts any object with choice/integers/random/shuffle as its random source
(CybORG/env.py:73-76, CybORG/Tests/utils.py:81-174 use exactly this seam).  TapeGenerator is such an
object: it draws from a seeded numpy PCG64 and logs each draw.

This module only works inside the build container (needs /root/reference); committed fixtures under
tests/golden/ are what travels.
"""
from __future__ import annotations

import os
import sys
from typing import Any

import numpy as np

REF_ROOT = os.environ.get("CC4_REFERENCE_ROOT", "/root/reference")
SHIM_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "refshim")


def _ensure_path():
    here = os.path.dirname(os.path.abspath(__file__))
    if here not in sys.path:
        sys.path.insert(0, here)
    for p in (SHIM_ROOT, REF_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    # shim must precede reference so gym/gymnasium resolve to the stand-ins
    sys.path.remove(SHIM_ROOT)
    sys.path.insert(0, SHIM_ROOT)


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "CybORG"))


# ------------------------------------------------------------------------------------------------
# Draw-site table: (file basename, line) -> tag.  Tags are shared with oracle/cc4_oracle.h (enum Site).
# Sites mapped to None are draws whose VALUE never influences the tracked state (IP allocation, OS
# distribution, ephemeral ports, the bandwidth shuffle, single-option blue session picks, ...): they are
# dropped from the tape.
# ------------------------------------------------------------------------------------------------
SITE = {
    # --- reset: EnterpriseScenarioGenerator
    ("EnterpriseScenarioGenerator.py", 263): None,   # subnet cidr pick
    ("EnterpriseScenarioGenerator.py", 335): None,   # internet host ip
    ("EnterpriseScenarioGenerator.py", 344): None,   # router ip
    ("EnterpriseScenarioGenerator.py", 350): 1,      # R_NUSER
    ("EnterpriseScenarioGenerator.py", 354): None,   # user ip
    ("EnterpriseScenarioGenerator.py", 360): 2,      # R_NSERVER
    ("EnterpriseScenarioGenerator.py", 494): None,   # distro
    ("EnterpriseScenarioGenerator.py", 558): 3,      # R_NADDON
    ("EnterpriseScenarioGenerator.py", 560): 4,      # R_ADDON
    ("EnterpriseScenarioGenerator.py", 575): 5,      # R_PID
    ("EnterpriseScenarioGenerator.py", 622): None,   # vuln roll (dead branch)
    ("EnterpriseScenarioGenerator.py", 654): None,   # blue start subnet (unused)
    ("EnterpriseScenarioGenerator.py", 668): 6,      # R_BLUE_PARENT
    ("EnterpriseScenarioGenerator.py", 781): 7,      # R_RED_SUBNET
    ("EnterpriseScenarioGenerator.py", 783): 8,      # R_RED_HOST
    ("EnterpriseScenarioGenerator.py", 804): None,   # red agent type (1 option)
    # --- host
    ("Host.py", 183): None,                          # ephemeral port
    ("Host.py", 185): None,                          # ephemeral port redraw
    ("Host.py", 200): 9,                             # CREATE_PID
    # --- controller
    ("SimulationController.py", 418): None,          # bandwidth shuffle
    # --- green
    ("EnterpriseGreenAgent.py", 60): 10,             # GREEN_ACT
    ("GreenLocalWork.py", 101): 11,                  # LW_SERVICE
    ("GreenLocalWork.py", 102): 12,                  # LW_RELIAB
    ("GreenLocalWork.py", 112): 13,                  # LW_FP
    ("GreenLocalWork.py", 118): 14,                  # LW_PHISH
    ("PhishingEmail.py", 88): 15,                    # PHISH_AGENT
    ("GreenAccessService.py", 115): 16,              # AS_DEST
    ("GreenAccessService.py", 207): 17,              # AS_FP
    # --- red policy
    ("FiniteStateRedAgent.py", 291): 18,             # FSM_HOST
    ("FiniteStateRedAgent.py", 311): 19,             # FSM_ACTION (choice with p)
    ("FiniteStateRedAgent.py", 330): 20,             # FSM_PARAM
    # --- DiscoveryFSRed only (host-state priorities + server bias, FiniteStateRedAgent._choose_host :252-293)
    ("FiniteStateRedAgent.py", 269): 32,             # FSMD_STATE_P   choice(states, p)
    ("FiniteStateRedAgent.py", 271): 33,             # FSMD_STATE_U   choice(states) when every priority is 0
    ("FiniteStateRedAgent.py", 278): 34,             # FSMD_ROLL      random() <= 0.75 -> a server
    ("FiniteStateRedAgent.py", 280): 35,             # FSMD_HOST_SRV
    ("FiniteStateRedAgent.py", 285): 36,             # FSMD_HOST_OTHER
    ("FiniteStateRedAgent.py", 287): 37,             # FSMD_HOST_SRV2 (only servers to pick from)
    ("FiniteStateRedAgent.py", 289): 38,             # FSMD_HOST_ALL  (no server among the options)
    # --- RandomSelectRedAgent only (RandomSelectRedAgent.get_action :37-92)
    ("RandomSelectRedAgent.py", 68): 39,             # RS_COMMAND     integers(0, len(list_commands))
    ("RandomSelectRedAgent.py", 81): 40,             # RS_PARAM       integers(0, len(param_opts)), only drawn for > 1 option
    # --- red actions
    ("Portscan.py", 49): 21,                         # PORTSCAN
    ("ExploitRemoteService.py", 61): 22,             # EXP_OTHER
    ("ExploitRemoteService.py", 65): 23,             # EXP_TOP
    ("ExploitAction.py", 83): 24,                    # EXP_SUCC
    ("ExploitAction.py", 102): 25,                   # EXP_DETECT
    ("SSHBruteForce.py", 44): 26,                    # SSH_HOP
    ("PrivilegeEscalate.py", 103): 27,               # PRIV_SESS
    ("DiscoverDeception.py", 89): 28,                # DD_TP
    ("DiscoverDeception.py", 95): 29,                # DD_FP
    ("RedSessionCheck.py", 37): 30,                  # RSC_PRIMARY
    # --- blue actions
    ("DecoyAction.py", 84): 31,                      # DECOY_FACTORY
    ("DecoyAction.py", 64): None,                    # blue session pick (1 option)
    ("Analyse.py", 62): None,
    ("Remove.py", 64): None,
    ("Restore.py", 62): None,
}

KIND_INT, KIND_F64 = 0, 1
ACTOR_RESET = 255
N_BLUE, N_RED, MAX_GREEN = 5, 6, 80
ACTOR_BLUE0, ACTOR_RED0, ACTOR_GREEN0 = 0, 5, 11
N_ACTORS = ACTOR_GREEN0 + MAX_GREEN  # 91


def actor_of(agent_name: str) -> int:
    kind, _, idx = agent_name.rpartition("_")
    idx = int(idx)
    if kind == "blue_agent":
        return ACTOR_BLUE0 + idx
    if kind == "red_agent":
        return ACTOR_RED0 + idx
    if kind == "green_agent":
        return ACTOR_GREEN0 + idx
    raise ValueError(agent_name)


class TapeGenerator:
    """Duck-typed numpy Generator replacement that records every draw.

    choice(a)        -> a[integers(0, len(a))]                       (1 integer draw)
    choice(a, p=p)   -> a[searchsorted(cumsum(p)/sum(p), random(), 'right')]  (1 double draw; numpy's rule)
    integers(lo, hi) -> lo + integers(0, hi-lo)
    random()         -> double in [0,1)
    shuffle(x)       -> Fisher-Yates, recorded draw by draw (site maps to None -> dropped)
    """

    def __init__(self, seed: int):
        self._g = np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))
        self.records: list[tuple[int, int, int, int, int, float]] = []  # (step, actor, site, kind, n, value)
        self.step = -1          # -1 = reset segment
        self.unknown_sites: set = set()

    def begin_segment(self, step: int):
        """Called by the harness before the final reset (step = -1) and before every tick."""
        self.step = step

    # -- value source: one integer in [0, n) / one double in [0, 1) for the draw at `site` by `actor` ---------------
    def _draw_int(self, site, actor, n: int) -> int:
        return int(self._g.integers(0, n))

    def _draw_f64(self, site, actor) -> float:
        return float(self._g.random())

    # -- context -------------------------------------------------------------------------------
    def _context(self):
        f = sys._getframe(2)
        site = (os.path.basename(f.f_code.co_filename), f.f_lineno)
        actor = ACTOR_RESET
        g = f
        while g is not None:
            if g.f_code.co_name == "step" and g.f_code.co_filename.endswith("SimulationController.py"):
                name = g.f_locals.get("agent_name")
                if name is not None:
                    actor = actor_of(name)
                break
            g = g.f_back
        return site, actor

    def _log(self, site, actor, kind, n, value):
        if site not in SITE:
            self.unknown_sites.add(site)
            return
        tag = SITE[site]
        if tag is None:
            return
        self.records.append((self.step, actor, tag, kind, n, value))

    # -- Generator API subset used by the reference ------------------------------------------------
    def random(self):
        site, actor = self._context()
        u = self._draw_f64(site, actor)
        self._log(site, actor, KIND_F64, 0, u)
        return u

    def integers(self, low, high=None, size=None, dtype=np.int64, endpoint=False):
        assert size is None
        site, actor = self._context()
        if high is None:
            low, high = 0, low
        if endpoint:
            high = high + 1
        n = int(high) - int(low)
        v = self._draw_int(site, actor, n)
        self._log(site, actor, KIND_INT, n, v)
        return int(low) + v

    def choice(self, a, size=None, replace=True, p=None, axis=0, shuffle=True):
        assert size is None
        site, actor = self._context()
        if isinstance(a, (int, np.integer)):
            n, seq = int(a), None
        else:
            seq = list(a)
            n = len(seq)
        if n == 0:
            raise ValueError("a cannot be empty unless no samples are taken")
        if p is None:
            idx = self._draw_int(site, actor, n)
            self._log(site, actor, KIND_INT, n, idx)
        else:
            p = np.asarray(p, dtype=np.float64)
            if abs(p.sum() - 1.0) > 1e-8:
                raise ValueError("probabilities do not sum to 1")
            cdf = p.cumsum()
            cdf /= cdf[-1]
            u = self._draw_f64(site, actor)
            idx = int(cdf.searchsorted(u, side="right"))
            self._log(site, actor, KIND_F64, n, u)
        return idx if seq is None else seq[idx]

    def shuffle(self, x):
        site, actor = self._context()
        for i in range(len(x) - 1, 0, -1):
            j = int(self._g.integers(0, i + 1))
            x[i], x[j] = x[j], x[i]
        self._log(site, actor, KIND_INT, len(x), 0)

    # gym spaces sometimes poke these
    def __getattr__(self, item):
        raise AttributeError(f"TapeGenerator has no '{item}' (reference used an RNG method the tape does not model)")


def philox4x32_10(ctr, k0: int, k1: int):
    """Philox4x32-10 (Salmon et al. 2011), the block function of the engine's production RNG (cc4_oracle.c philox4x32_10)."""
    c0, c1, c2, c3 = (int(x) & 0xFFFFFFFF for x in ctr)
    k0 &= 0xFFFFFFFF
    k1 &= 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = 0xD2511F53 * c0, 0xCD9E8D57 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & 0xFFFFFFFF, p1 & 0xFFFFFFFF, ((p0 >> 32) ^ c3 ^ k1) & 0xFFFFFFFF, p0 & 0xFFFFFFFF
        k0, k1 = (k0 + 0x9E3779B9) & 0xFFFFFFFF, (k1 + 0xBB67AE85) & 0xFFFFFFFF
    return c0, c1, c2, c3


SITE_FSM_PARAM, SITE_LW_SERVICE, SITE_DD_TP, SITE_DD_FP = 20, 11, 28, 29


class PhiloxSeamGenerator(TapeGenerator):
    """The engine's PRODUCTION random source, served to the live reference through the same seam (env.py:73-76).

    Every mapped draw site gets the value the engine would draw at that point: stream (actor, tick) of episode
    (seed, episode) is the sequence of 32-bit words of Philox4x32-10 blocks with key (seed lo, seed hi ^ episode hi) and
    counter (word index / 4, actor, tick + 1 [0 for the reset stream, actor 91], episode lo) — cc4_oracle.c rng_word,
    cc4_device.cuh philox_block.  Word rules (cc4_oracle.c rng_int / rng_int_more / rng_f64, exec_red AC_DD):
      * integer in [0, n) = high half of word * n; double = word / 2^32; one word per draw;
      * a one-option FSM_PARAM pick spends no word;
      * LW_SERVICE is cut from the low half of the word the actor's GREEN_ACT draw used;
      * DiscoverDeception owns two words per process (DD_TP, DD_FP) whether or not the reference looks at the second.
    Draw sites the engine does not model (SITE -> None: addresses, ephemeral ports, the bandwidth shuffle, ...) are
    served from a side PCG64 and never touch the Philox streams.  A reference episode driven by this generator must
    therefore equal, tick for tick, the oracle and the CUDA engine running in rng="philox" mode with the same
    (seed, episode): tests/golden/px_*.npz.
    """

    NAG = N_ACTORS      # the reset stream's actor id (cc4_oracle.c: actor == NAG)

    def __init__(self, seed: int, episode: int):
        super().__init__(seed ^ 0x5EA3)          # side stream for the unmapped sites
        self.seed, self.episode = int(seed), int(episode)
        self._k0 = self.seed & 0xFFFFFFFF
        self._k1 = ((self.seed >> 32) ^ (self.episode >> 32)) & 0xFFFFFFFF
        self.begin_segment(-1)

    def begin_segment(self, step: int):
        self.step = step
        self._ctr: dict[int, int] = {}          # actor -> next word index of its stream for this tick
        self._frac: dict[int, int] = {}         # actor -> low half of its last integer draw's product
        self._dd_second: dict[int, int] = {}    # actor -> the reserved second word of the current DiscoverDeception process

    def _word(self, actor: int) -> int:
        a = self.NAG if actor == ACTOR_RESET else actor
        j = self._ctr.get(a, 0)
        self._ctr[a] = j + 1
        tick = 0 if a == self.NAG else self.step + 1
        return philox4x32_10((j >> 2, a, tick, self.episode & 0xFFFFFFFF), self._k0, self._k1)[j & 3]

    def _mapped(self, site) -> int | None:
        return SITE.get(site)

    def _draw_int(self, site, actor, n: int) -> int:
        tag = self._mapped(site)
        if tag is None:
            return int(self._g.integers(0, n))
        a = self.NAG if actor == ACTOR_RESET else actor
        if tag == SITE_FSM_PARAM and n == 1:
            return 0
        if tag == SITE_LW_SERVICE:
            p = self._frac[a] * n
        else:
            p = self._word(actor) * n
        self._frac[a] = p & 0xFFFFFFFF
        return p >> 32

    def _draw_f64(self, site, actor) -> float:
        tag = self._mapped(site)
        if tag is None:
            return float(self._g.random())
        a = self.NAG if actor == ACTOR_RESET else actor
        if tag == SITE_DD_TP:
            w = self._word(actor)
            self._dd_second[a] = self._word(actor)
        elif tag == SITE_DD_FP:
            w = self._dd_second.pop(a)
        else:
            w = self._word(actor)
        return w / 4294967296.0

    def shuffle(self, x):
        site, actor = self._context()
        assert SITE.get(site) is None, "a modelled shuffle would need a word rule"
        for i in range(len(x) - 1, 0, -1):
            j = int(self._g.integers(0, i + 1))
            x[i], x[j] = x[j], x[i]


# ------------------------------------------------------------------------------------------------
# Canonical state extraction (layout documented in oracle/canon.py)
# ------------------------------------------------------------------------------------------------
from canon import (SUBNETS_CREATION, SUBNET_IDX, PK, SERVICE_NAMES, FSM_STATES, PORT_BIT, ACTION_CODE,  # noqa: E402
                   host_slot, encode)


def _service_code(name) -> int:
    s = getattr(name, "name", None) or str(name)
    return SERVICE_NAMES.index(s)


def _proc_kind(proc) -> int:
    name = getattr(proc.name, "name", None) or str(proc.name)
    if proc.decoy_type.name == "EXPLOIT":
        return {"apache2": PK["D_APACHE"], "Tomcat.exe": PK["D_TOMCAT"], "haraka": PK["D_HARAKA"],
                "vsftpd": PK["D_VSFTPD"]}[name]
    if name in ("SSHD", "OTSERVICE", "APACHE2", "MYSQLD", "SMTP"):
        return {"SSHD": 0, "OTSERVICE": 1, "APACHE2": 2, "MYSQLD": 3, "SMTP": 4}[name]
    if name.startswith("cmd."):
        return PK["SHELL"]
    if name == "RED_ABSTRACT_SESSION":
        return PK["REDABS"]
    return PK["SESSION"]


def _inprog(ctrl, name):
    ip = ctrl.actions_in_progress.get(name)
    if ip is None:
        return (-1, 0)
    return (ACTION_CODE[type(ip["action"]).__name__], int(ip["remaining_ticks"]))


class RefEpisode:
    """One reference environment driven through BlueFlatWrapper with index actions."""

    def __init__(self, seed: int, steps: int = 500, red: str = "fsm", rng: str = "pcg64", episode: int = 0, pad_spaces: bool = False):
        _ensure_path()
        from CybORG import CybORG
        from CybORG.Agents import SleepAgent, EnterpriseGreenAgent, FiniteStateRedAgent
        if red == "discovery":      # CybORG/Agents/SimpleAgents/FSMRedVariants.py:72-122
            from CybORG.Agents.SimpleAgents.FSMRedVariants import DiscoveryFSRed as FiniteStateRedAgent
        if red == "random_select":  # CybORG/Agents/SimpleAgents/RandomSelectRedAgent.py:8-139 (the only agent that uses Withdraw)
            from CybORG.Agents.SimpleAgents.RandomSelectRedAgent import RandomSelectRedAgent as FiniteStateRedAgent
        from CybORG.Agents.Wrappers import BlueFlatWrapper
        from CybORG.Simulator.Scenarios import EnterpriseScenarioGenerator

        self.tape = TapeGenerator(seed) if rng == "pcg64" else PhiloxSeamGenerator(seed, episode)
        sg = EnterpriseScenarioGenerator(blue_agent_class=SleepAgent, green_agent_class=EnterpriseGreenAgent,
                                         red_agent_class=FiniteStateRedAgent, steps=steps)
        # construction runs create_scenario once; wrapper.reset() runs it again. Only the LAST scenario
        # matters; the tape keeps only the draws made after the final reset started.
        self.cyborg = CybORG(sg, seed=self.tape)
        self.env = BlueFlatWrapper(self.cyborg, pad_spaces=pad_spaces)
        self.tape.records.clear()
        self.tape.begin_segment(-1)
        self.obs, self.info = self.env.reset()
        self.ctrl = self.cyborg.environment_controller
        self.t = 0

    def step(self, actions, messages=None):
        if isinstance(actions, (list, tuple, np.ndarray)):
            actions = {f"blue_agent_{i}": int(a) for i, a in enumerate(actions)}
        if isinstance(messages, (list, tuple, np.ndarray)):
            messages = {f"blue_agent_{i}": np.array([(int(m) >> b) & 1 for b in range(8)], dtype=bool)
                        for i, m in enumerate(messages)}
        self.tape.begin_segment(self.t)
        obs, rew, term, trunc, info = self.env.step(actions, messages=messages)
        self.t += 1
        self.obs = obs
        return obs, rew, term, trunc, info

    def obs_array(self) -> np.ndarray:
        out = np.zeros((5, 210), dtype=np.uint8)
        for i in range(5):
            v = np.asarray(self.obs[f"blue_agent_{i}"])
            out[i, :len(v)] = v
        return out

    def mask_array(self) -> np.ndarray:
        out = np.zeros((5, 242), dtype=np.uint8)
        for i in range(5):
            m = self.env.action_mask(f"blue_agent_{i}")
            out[i, :len(m)] = m
        return out

    def state(self) -> dict[str, Any]:
        import ipaddress
        st = self.ctrl.state
        out: dict[str, Any] = {"phase": int(st.mission_phase), "step": int(self.ctrl.step_count)}
        blocks = [0] * 9
        for to, froms in st.blocks.items():
            for fr in froms:
                blocks[SUBNET_IDX[to]] |= 1 << SUBNET_IDX[fr]
        out["blocks"] = blocks
        hosts = {}
        for hname, host in st.hosts.items():
            procs = [(int(p.pid), _proc_kind(p), int(p.user == "root")) for p in host.processes]
            svcs = [(_service_code(n), int(sv.process), int(bool(sv.active)), int(sv.get_service_reliability()))
                    for n, sv in host.services.items()]
            ev = (int(len(host.events.process_creation) > 0), int(len(host.events.network_connections) > 0),
                  int(len(host.events.old_process_creation) > 0), int(len(host.events.old_network_connections) > 0))
            hosts[host_slot(hname)] = dict(procs=procs, svcs=svcs, ev=ev)
        out["hosts"] = hosts
        red = []
        for r in range(6):
            name = f"red_agent_{r}"
            ai = self.ctrl.agent_interfaces[name]
            sess = [(int(sid), host_slot(s.hostname), int(type(s).__name__ == "RedAbstractSession"),
                     int(s.username == "root"), int(s.pid)) for sid, s in st.sessions[name].items()]
            asp = ai.action_space
            ipk = sorted(host_slot(st.ip_addresses[ip]) for ip, v in asp.ip_address.items() if v)
            hnk = sorted(host_slot(h) for h, v in asp.hostname.items() if v)
            snk = 0
            for c, v in asp.subnet.items():
                if v:
                    snk |= 1 << SUBNET_IDX[st.subnets_cidr_to_name[c]]
            ssk = [int(k) for k, v in asp.server_session.items() if v]
            fsm = []
            for ip, hs in self._fsm_view(name).items():
                slot = -1 if ip is None else host_slot(st.ip_addresses[ipaddress.IPv4Address(ip)])
                fsm.append((slot, FSM_STATES.index(hs["state"]), int(hs["hostname"] is not None)))
            ports = []
            for sid, s in st.sessions[name].items():
                for ip, pl in getattr(s, "ports", {}).items():
                    mask = 0
                    for p in pl:
                        mask |= 1 << PORT_BIT.get(int(p), 7)
                    ports.append((int(sid), host_slot(st.ip_addresses[ip]), mask))
            red.append(dict(active=int(ai.active), count=int(st.sessions_count[name]), sessions=sess, ip_known=ipk,
                            host_known=hnk, subnet_known=snk, server_sessions=ssk, fsm=fsm, ports=sorted(ports),
                            inprog=_inprog(self.ctrl, name)))
        out["red"] = red
        blue = []
        for b in range(5):
            name = f"blue_agent_{b}"
            sus = []
            for s in st.sessions[name].values():
                for h, pids in getattr(s, "sus_pids", {}).items():
                    sus += [(host_slot(h), int(p)) for p in pids]
            sus.sort(key=lambda x: x[0])
            blue.append(dict(sus=sus, inprog=_inprog(self.ctrl, name)))
        out["blue"] = blue
        return out

    def _fsm_view(self, name: str) -> dict:
        """FiniteStateRedAgent.host_states as it will be right after the agent digests this tick's observation
        (the first thing get_action does next tick, FiniteStateRedAgent.py:91-93), computed with the reference's
        own methods on a copy.  Inactive agents are not asked for an action (AgentInterface.py:133-136)."""
        import copy
        ai = self.ctrl.agent_interfaces[name]
        agent = ai.agent
        if not hasattr(agent, "host_states"):
            return {}
        if not ai.active:
            return agent.host_states
        twin = copy.copy(agent)
        twin.host_states = copy.deepcopy(agent.host_states)
        twin.host_service_decoy_status = copy.deepcopy(agent.host_service_decoy_status)
        obs = copy.deepcopy(self.ctrl.get_last_observation(name).data)
        success = obs.pop("success", None)
        action = obs.pop("action", None)
        twin._host_state_transition(action, success)
        twin._process_new_observations(obs)
        twin._session_removal_state_change(obs)
        return twin.host_states

    def canon(self) -> list[int]:
        return encode(self.state())


# ------------------------------------------------------------------------------------------------
# Fixture generation
# ------------------------------------------------------------------------------------------------
SLEEP_INDEX = (49, 49, 49, 49, 145)  # BlueFixedActionWrapper._populate_action_space: Analyse*16n, Monitor, Remove*16n, Restore*16n, Sleep
N_ACTIONS = (82, 82, 82, 82, 242)


def blue_policy(kind: str, rng: np.random.Generator, mask: np.ndarray, obs: np.ndarray | None = None):
    """Synthetic blue inputs (NOT drawn from the environment RNG)."""
    if kind in ("reactive", "lockdown"):
        # reactive: answer alerts in the observation vector (BlueFlatWrapper.py:172-256: per own subnet 27 policy bits,
        # 16 malicious-process flags, 16 network-connection flags) with Remove / Restore / Analyse on the flagged host,
        # otherwise decoys and the odd traffic-control action.  lockdown: the same with frequent Block/Allow.
        acts = []
        for b in range(5):
            n = 3 if b == 4 else 1
            H, T = 16 * n, 8 * n
            flagged = [i * 16 + h for i in range(n) for h in range(16)
                       if obs[b, 1 + 59 * i + 27 + h] or obs[b, 1 + 59 * i + 43 + h]]
            u = rng.random()
            p_traffic = 0.35 if kind == "lockdown" else 0.08
            if flagged and u < 0.7:
                h = int(rng.choice(flagged))
                v = rng.random()
                a = (H + 1 + h) if v < 0.5 else (2 * H + 1 + h) if v < 0.8 else h     # Remove / Restore / Analyse
            elif u < 0.7 + p_traffic:
                a = 3 * H + 2 + int(rng.integers(0, 2 * T))                            # Allow / Block
            elif u < 0.9:
                a = 3 * H + 2 + 2 * T + int(rng.integers(0, H))                        # DeployDecoy
            else:
                a = H if rng.random() < 0.5 else 3 * H + 1                             # Monitor / Sleep
            acts.append(a)
        return acts, [int(rng.integers(0, 256)) for _ in range(5)]
    if kind == "sleep":
        return list(SLEEP_INDEX), [0] * 5
    if kind == "random":          # uniform over the whole index range (masked-out entries act as Sleep)
        return [int(rng.integers(0, n)) for n in N_ACTIONS], [int(rng.integers(0, 256)) for _ in range(5)]
    if kind == "masked":          # uniform over valid entries only
        return [int(rng.choice(np.flatnonzero(mask[i, :N_ACTIONS[i]]))) for i in range(5)], \
               [int(rng.integers(0, 256)) for _ in range(5)]
    if kind == "decoy":           # every agent deploys a decoy on server_host_0 of its first subnet, every tick (process-list growth)
        return [66, 66, 66, 66, 194], [0] * 5
    raise ValueError(kind)


def run_episode(seed: int, policy: str = "sleep", steps: int = 500, n_ticks: int | None = None,
                with_state: bool = True, red: str = "fsm", rng: str = "pcg64", episode: int = 0) -> dict[str, np.ndarray]:
    """Run one reference episode and return the fixture arrays.  rng="philox": the reference draws the engine's
    production-mode values for (seed, episode) (PhiloxSeamGenerator) instead of a seeded PCG64."""
    ep = RefEpisode(seed, steps, red, rng, episode)
    n_ticks = steps - 1 if n_ticks is None else n_ticks
    prng = np.random.Generator(np.random.PCG64(seed + 7919))
    mask = ep.mask_array()
    obs = [ep.obs_array()]
    canon = [ep.canon()] if with_state else []
    acts, msgs, rews, dones = [], [], [], []
    for _ in range(n_ticks):
        a, m = blue_policy(policy, prng, mask, obs[-1])
        o, r, term, trunc, _ = ep.step(a, m)
        acts.append(a)
        msgs.append(m)
        rews.append(r["blue_agent_0"])
        dones.append(int(term["blue_agent_0"]))
        obs.append(ep.obs_array())
        if with_state:
            canon.append(ep.canon())
    if ep.tape.unknown_sites:
        raise RuntimeError(f"unmapped RNG call sites: {sorted(ep.tape.unknown_sites)}")
    rec = ep.tape.records
    t_val = np.zeros(len(rec), dtype=np.uint64)
    for i, x in enumerate(rec):
        t_val[i] = np.float64(x[5]).view(np.uint64) if x[3] == KIND_F64 else np.uint64(x[5])
    off = np.zeros(len(canon) + 1, dtype=np.int64)
    for i, c in enumerate(canon):
        off[i + 1] = off[i] + len(c)
    return dict(
        seed=np.int64(seed), steps=np.int32(steps), policy=np.array(policy), red=np.array(red),
        rng=np.array(rng), episode=np.int64(episode),
        t_step=np.array([x[0] for x in rec], dtype=np.int32), t_actor=np.array([x[1] for x in rec], dtype=np.uint8),
        t_site=np.array([x[2] for x in rec], dtype=np.uint8), t_kind=np.array([x[3] for x in rec], dtype=np.uint8),
        t_n=np.array([x[4] for x in rec], dtype=np.uint16), t_val=t_val,
        actions=np.array(acts, dtype=np.int16).reshape(-1, 5), msgs=np.array(msgs, dtype=np.uint8).reshape(-1, 5),
        obs=np.stack(obs), reward=np.array(rews, dtype=np.float32), done=np.array(dones, dtype=np.uint8), mask=mask,
        canon=np.array([v for c in canon for v in c], dtype=np.int32), canon_off=off,
    )


def run_padded_episode(seed: int, steps: int, n_ticks: int, episode: int = 0) -> dict[str, np.ndarray]:
    """BlueFlatWrapper(pad_spaces=True) of the live reference on the Philox seam (BlueFixedActionWrapper.py:311-328,
    BlueFlatWrapper.py:251-254): every agent sees 210 observation entries and 242 actions; indices are drawn over the whole
    padded range, so the `[Padding] Sleep` entries are exercised."""
    import json
    ep = RefEpisode(seed, steps, "fsm", "philox", episode, pad_spaces=True)
    prng = np.random.Generator(np.random.PCG64(seed + 7919))
    mask = ep.mask_array()
    obs, acts, rews = [ep.obs_array()], [], []
    for _ in range(n_ticks):
        a = [int(prng.integers(0, 242)) for _ in range(5)]
        _, r, _, _, _ = ep.step(a, [0] * 5)
        acts.append(a)
        rews.append(r["blue_agent_0"])
        obs.append(ep.obs_array())
    labels = [ep.env.action_labels(f"blue_agent_{b}") for b in range(5)]
    spaces = [[int(ep.env.action_space(f"blue_agent_{b}").n), len(ep.env.observation_space(f"blue_agent_{b}"))] for b in range(5)]
    return dict(seed=np.int64(seed), episode=np.int64(episode), steps=np.int32(steps), rng=np.array("philox"),
                actions=np.array(acts, dtype=np.int16), obs=np.stack(obs), reward=np.array(rews, dtype=np.float32), mask=mask,
                labels=np.array(json.dumps(labels)), spaces=np.array(spaces, dtype=np.int32))


def dict_obs_summary(obs: dict) -> dict:
    """The reconstructible part of a blue agent's dict observation (env.py:245): success, the action as it prints, the hosts
    that carry an entry and, per host, the pids of its `Processes` entries (process-creation events, decoys)."""
    hosts = {}
    for k, v in obs.items():
        if k in ("success", "action", "message"):
            continue
        hosts[str(k)] = sorted(int(p["PID"]) for p in v.get("Processes", []) if "PID" in p)
    return {"success": obs["success"].name, "action": str(obs.get("action")), "hosts": hosts}


def run_dictobs_episode(seed: int, policy: str, steps: int, episode: int = 0) -> dict[str, np.ndarray]:
    """A reference episode on the Philox seam (so that the facade's production mode with the same seed IS this episode) with
    the per-tick summaries of the five blue dict observations: fixture for cage_challenge_4_b200/observations.py."""
    import json
    ep = RefEpisode(seed, steps, "fsm", "philox", episode)
    prng = np.random.Generator(np.random.PCG64(seed + 7919))
    mask = ep.mask_array()
    acts, msgs, summ, rews = [], [], [], []
    for _ in range(steps - 1):
        a, m = blue_policy(policy, prng, mask, ep.obs_array())
        _, r, _, _, _ = ep.step(a, m)
        acts.append(a)
        msgs.append(m)
        rews.append(r["blue_agent_0"])
        summ.append([dict_obs_summary(ep.cyborg.get_observation(f"blue_agent_{b}")) for b in range(5)])
    return dict(seed=np.int64(seed), episode=np.int64(episode), steps=np.int32(steps), policy=np.array(policy), rng=np.array("philox"),
                actions=np.array(acts, dtype=np.int16), msgs=np.array(msgs, dtype=np.uint8), reward=np.array(rews, dtype=np.float32),
                mask=mask, summaries=np.array(json.dumps(summ)))


if __name__ == "__main__":
    import argparse
    import time
    ap = argparse.ArgumentParser(description="generate golden fixtures from the live reference")
    ap.add_argument("--out", default=os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden"))
    ap.add_argument("--spec", nargs="*", default=["0:sleep:500", "1:sleep:500", "2:random:500", "3:random:500",
                                                  "4:masked:500", "5:masked:500"])
    ap.add_argument("--padded", nargs="*", default=[], help="seed:steps:ticks fixtures of BlueFlatWrapper(pad_spaces=True) (Philox seam, episode 0)")
    ap.add_argument("--dictobs", nargs="*", default=[], help="seed:policy:steps fixtures of the blue dict observations (Philox seam, episode 0)")
    args = ap.parse_args()
    os.makedirs(args.out, exist_ok=True)
    for spec in args.dictobs:
        seed, policy, steps = spec.split(":")
        fx = run_dictobs_episode(int(seed), policy, int(steps))
        path = os.path.join(args.out, f"dictobs_s{seed}_{policy}_{steps}.npz")
        np.savez_compressed(path, **fx)
        print(f"{path}: {len(fx['actions'])} ticks, {os.path.getsize(path) / 1024:.0f} KiB")
    for spec in args.padded:
        seed, steps, ticks = (int(x) for x in spec.split(":"))
        fx = run_padded_episode(seed, steps, ticks)
        path = os.path.join(args.out, f"pad_s{seed}_{steps}_{ticks}.npz")
        np.savez_compressed(path, **fx)
        print(f"{path}: {ticks} ticks, {os.path.getsize(path) / 1024:.0f} KiB")
    if args.dictobs or args.padded:
        sys.exit(0)
    for spec in args.spec:
        seed, policy, steps, *rest = spec.split(":")      # seed:policy:steps[:red[:episode]]  (red = fsm | discovery | random_select; episode => Philox seam)
        red = rest[0] if rest else "fsm"
        t0 = time.time()
        if len(rest) > 1:       # production-mode fixture: the reference driven by the engine's Philox streams of (seed, episode)
            fx = run_episode(int(seed), policy, int(steps), red=red, rng="philox", episode=int(rest[1]))
            path = os.path.join(args.out, f"px_s{seed}_e{rest[1]}_{policy}_{steps}{'' if red == 'fsm' else '_' + red}.npz")
        else:
            fx = run_episode(int(seed), policy, int(steps), red=red)
            path = os.path.join(args.out, f"ep_s{seed}_{policy}_{steps}{'' if red == 'fsm' else '_' + red}.npz")
        np.savez_compressed(path, **fx)
        print(f"{path}: {len(fx['t_val'])} draws, {len(fx['canon'])} state ints, return {fx['reward'].sum():.1f}, "
              f"{time.time() - t0:.1f}s, {os.path.getsize(path) / 1024:.0f} KiB")
