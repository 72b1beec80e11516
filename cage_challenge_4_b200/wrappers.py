"""Host-side mirror of the reference's environment / wrapper API for the hot path, backed by the CUDA engine.

Same names, argument meaning and return structure as the reference classes, so code written against
`CybORG(...)` + `BlueFlatWrapper` / `BlueEnterpriseWrapper` / `EnterpriseMAE` (reference: CybORG/env.py:53,
CybORG/Agents/Wrappers/BlueFixedActionWrapper.py:44, BlueFlatWrapper.py:40, BlueEnterpriseWrapper.py:13,
EnterpriseMAE.py:10) — in particular the loop of CybORG/Evaluation/evaluation.py:56-134 — runs unchanged on it.
Each wrapper object presents ONE episode of a `BatchedCC4` batch (index `env_index`); the batched API itself is
`cage_challenge_4_b200.BatchedCC4`.

What is NOT mirrored (out of the hot path): executable `Action` objects (actions are the wrapper's
integer indices; `BlueFixedActionWrapper.actions()` returns descriptors that print like the reference's objects) and
user-supplied red / green agent classes (the engine scripts FiniteStateRedAgent, DiscoveryFSRed, RandomSelectRedAgent and EnterpriseGreenAgent;
anything else raises).  The engine's state has no IP addresses (in the reference they are random labels that never
influence a transition): `get_ip_map` / `get_cidr_map` and the labels use a fixed synthetic numbering (10.0.<subnet>.0/24,
host k of a subnet at .k+2) instead.  The per-episode classes present ONE episode and therefore require `num_envs == 1`;
batches are stepped through `BatchedCC4` (engine.py) and `evaluation.run_evaluation_batched`.
Seeds select Philox streams, not numpy PCG64 streams, so a seeded trajectory is reproducible but differs from the
reference's trajectory for the same integer (parity is established through RNG-tape replay, tests/).
"""
from __future__ import annotations

import enum
import functools
import ipaddress
import itertools
from typing import Any

import numpy as np
import torch

from .engine import BatchedCC4, N_ACTIONS, OBS_SIZES, SLEEP_INDEX
from .spaces import Discrete, MultiBinary, MultiDiscrete

SUBNETS_CREATION = [
    "restricted_zone_a_subnet", "operational_zone_a_subnet", "restricted_zone_b_subnet", "operational_zone_b_subnet",
    "contractor_network_subnet", "public_access_zone_subnet", "admin_network_subnet", "office_network_subnet", "internet_subnet",
]
SUBNETS_SORTED = sorted(SUBNETS_CREATION)
BLUE_SUBNETS = [  # EnterpriseScenarioGenerator.py:643-649
    ["restricted_zone_a_subnet"], ["operational_zone_a_subnet"], ["restricted_zone_b_subnet"], ["operational_zone_b_subnet"],
    ["public_access_zone_subnet", "admin_network_subnet", "office_network_subnet"],
]
NUM_SUBNETS, MAX_USER_HOSTS, MAX_SERVER_HOSTS, MESSAGE_LENGTH, NUM_MESSAGES = 9, 10, 6, 8, 4
MAX_HOSTS = MAX_USER_HOSTS + MAX_SERVER_HOSTS
EMPTY_MESSAGE = np.zeros(MESSAGE_LENGTH, dtype=bool)


# -- marker classes with the reference's names (CybORG/Agents/SimpleAgents) ---------------------------------------------
class BaseAgent:
    def __init__(self, name: str = None, np_random=None):
        self.name = name

    def get_action(self, observation, action_space):
        raise NotImplementedError


class SleepAgent(BaseAgent):
    """ConstantAgent.py:32 — the default blue placeholder; blue actions come from the wrapper's step()."""


class EnterpriseGreenAgent(BaseAgent):
    """Scripted in-kernel (EnterpriseGreenAgent.py:42-78)."""


class FiniteStateRedAgent(BaseAgent):
    """Scripted in-kernel (FiniteStateRedAgent.py:58-336)."""


class DiscoveryFSRed(FiniteStateRedAgent):
    """Scripted in-kernel (FSMRedVariants.py:72-122: host-state priorities, server bias, discovery-weighted actions)."""


class RandomSelectRedAgent(BaseAgent):
    """Scripted in-kernel (RandomSelectRedAgent.py:8-139: uniformly random valid command and parameters, Withdraw included)."""


class EnterpriseScenarioGenerator:
    """Same constructor as Simulator/Scenarios/EnterpriseScenarioGenerator.py:95-101."""
    MIN_USER_HOSTS, MAX_USER_HOSTS, MIN_SERVER_HOSTS, MAX_SERVER_HOSTS = 3, 10, 1, 6
    MAX_BANDWIDTH, MESSAGE_LENGTH = 100, 8

    def __init__(self, blue_agent_class=None, red_agent_class=None, green_agent_class=None, steps: int = 100):
        red_name = "FiniteStateRedAgent" if red_agent_class is None else getattr(red_agent_class, "__name__", "")
        if red_name not in ("FiniteStateRedAgent", "DiscoveryFSRed", "RandomSelectRedAgent"):
            raise NotImplementedError("the engine scripts FiniteStateRedAgent, DiscoveryFSRed and RandomSelectRedAgent; other red agent classes cannot run in the kernel")
        self.red_agent = red_name
        if green_agent_class not in (None, EnterpriseGreenAgent) and getattr(green_agent_class, "__name__", "") != "EnterpriseGreenAgent":
            raise NotImplementedError("the engine scripts EnterpriseGreenAgent only")
        self.blue_agent_class, self.red_agent_class, self.green_agent_class, self.steps = blue_agent_class, red_agent_class, green_agent_class, steps


class TernaryEnum(enum.Enum):
    """Shared/Enums.py:5-24 (every member is truthy, as in the reference)."""
    TRUE = 1
    UNKNOWN = 2
    FALSE = 3
    IN_PROGRESS = 4


class ActionDescriptor:
    """What `BlueFixedActionWrapper.actions()` lists in place of the reference's Action objects: the action class name and
    its parameters, printing like the reference object does (`Restore <hostname>`, `BlockTrafficZone`, `Sleep`:
    Analyse.py:73, Action.py:31).  The engine executes INDICES; a descriptor is a label with structure."""

    def __init__(self, name: str, **params):
        self.name, self.params = name, params
        for k, v in params.items():
            setattr(self, k, v)

    def get_params(self) -> dict:
        return dict(self.params)

    def __str__(self):
        return f"{self.name} {self.params['hostname']}" if "hostname" in self.params else self.name

    __repr__ = __str__

    def __eq__(self, other):
        return isinstance(other, ActionDescriptor) and (self.name, self.params) == (other.name, other.params)

    def __hash__(self):
        return hash((self.name, tuple(sorted(self.params.items()))))


class Results:
    """Shared/Results.py:9: what the single-agent `CybORG.step` returns (only these attributes are used in CC4)."""

    def __init__(self, observation=None, done=None, reward=None, info=None, action_space=None, action=None):
        self.observation, self.done, self.reward, self.info, self.action_space, self.action = observation, done, reward, info, action_space, action


class LastAction(int):
    """An action index that prints as the action's label."""

    def __new__(cls, index: int, label: str):
        self = super().__new__(cls, index)
        self.label = label
        return self

    def __repr__(self):
        return self.label

    __str__ = __repr__


# -- names and (synthetic) addresses ---------------------------------------------------------------------------------
def subnet_hosts(subnet: str, n_user: int = MAX_USER_HOSTS, n_server: int = MAX_SERVER_HOSTS) -> list[str]:
    """Hostnames of a subnet in creation order (EnterpriseScenarioGenerator.py:342-366): router, users, servers."""
    return [f"{subnet}_router"] + [f"{subnet}_user_host_{i}" for i in range(n_user)] + [f"{subnet}_server_host_{i}" for i in range(n_server)]


def subnet_cidr(subnet: str) -> ipaddress.IPv4Network:
    return ipaddress.IPv4Network(f"10.0.{SUBNETS_CREATION.index(subnet)}.0/24")


def host_ip(hostname: str) -> ipaddress.IPv4Address:
    """Synthetic address of a host: 10.0.<subnet>.1 router, .2 + k user_host_k, .12 + k server_host_k (the reference draws
    random addresses, EnterpriseScenarioGenerator.py:335-354; nothing ever depends on their values)."""
    if hostname == "root_internet_host_0":
        return ipaddress.IPv4Address("10.0.8.2")
    for i, sn in enumerate(SUBNETS_CREATION):
        if hostname.startswith(sn + "_"):
            rest = hostname[len(sn) + 1:]
            if rest == "router":
                return ipaddress.IPv4Address(f"10.0.{i}.1")
            kind, _, k = rest.rpartition("_")
            return ipaddress.IPv4Address(f"10.0.{i}.{(2 if kind == 'user_host' else 12) + int(k)}")
    raise KeyError(hostname)


class CybORG:
    """Facade with the constructor and the CC4-relevant methods of CybORG/env.py:53-460.  `agents` is accepted and ignored
    exactly as the evaluation script uses it (it passes "sim", evaluation.py:70).  Extra keywords choose the batch:
    num_envs episodes on `device`; the per-episode methods (parallel_step, get_observation, ...) act on a one-episode batch."""

    def __init__(self, scenario_generator: EnterpriseScenarioGenerator, agents=None, seed: int | None = None, *,
                 num_envs: int = 1, device: int = 0):
        assert isinstance(scenario_generator, EnterpriseScenarioGenerator), "scenario_generator must be an EnterpriseScenarioGenerator"
        self.scenario_generator = scenario_generator
        self.seed = int(np.random.SeedSequence().entropy % (2**63)) if seed is None else int(seed)
        self.engine = BatchedCC4(num_envs, device=device, seed=self.seed, steps=scenario_generator.steps, red_agent=scenario_generator.red_agent)
        self._episodes_started = 0
        self._last_actions = None
        self._label_source = None       # set by the wrapper: agent name -> list of action labels
        self._blue = [f"blue_agent_{i}" for i in range(5)]
        self._rewards = {"Blue": {"BlueRewardMachine": 0.0, "action_cost": 0.0}, "Red": {"None": 0.0, "action_cost": 0},
                         "Green": {"None": 0.0, "action_cost": 0}}
        self._done = False
        self._helper = None
        self._tracker = None            # observations.BlueObservationTracker once dict observations are wanted
        self._want_dict = False
        self._ticks = 0
        self.reset()

    # ---- the one place a tick is submitted (used by parallel_step and by the wrappers) ----------------------------------
    def _single(self, what: str):
        if self.engine.num_envs != 1:
            raise ValueError(f"{what} presents ONE episode but this CybORG holds a batch of {self.engine.num_envs}: a call would advance the other "
                             f"episodes with Sleep and drop their outputs.  Step batches through BatchedCC4 (cyborg.engine) or "
                             f"evaluation.run_evaluation_batched; use num_envs=1 with the per-episode API")

    def _tick(self, indices: dict[str, int], message_bytes: dict[str, int]):
        """One tick of the (single) episode with BlueFixedActionWrapper indices; returns (obs uint8 [5, 210], reward, done)."""
        eng = self.engine
        acts = torch.tensor([SLEEP_INDEX], dtype=torch.int32)
        msgs = torch.zeros((1, 5), dtype=torch.uint8)
        for a, v in indices.items():
            acts[0, int(a.rsplit("_", 1)[1])] = int(v)
        for a, m in message_bytes.items():
            msgs[0, int(a.rsplit("_", 1)[1])] = int(m)
        prev_row = self._obs[0].cpu().numpy() if self._tracker is not None else None
        obs, rew, done = eng.step(acts.to(eng.device), msgs.to(eng.device))
        self._last_actions = acts.numpy()
        self._obs = obs
        self._ticks += 1
        if self._tracker is not None:
            self._tracker.after_tick(prev_row, self._labels())
        r, d = float(rew[0]), bool(done[0])
        # Restore costs -1 on submission (Restore.py:70, SimulationController.py:310-311); everything else is BlueRewardMachine
        cost = -float(sum(1 for b in range(5) if self._is_restore(b, int(acts[0, b]))))
        self._rewards = {"Blue": {"BlueRewardMachine": r - cost, "action_cost": cost}, "Red": {"None": 0.0, "action_cost": 0},
                         "Green": {"None": 0.0, "action_cost": 0}}
        self._done = d
        return obs[0].cpu().numpy(), r, d

    def _is_restore(self, b: int, idx: int) -> bool:
        h = 48 if b == 4 else 16
        return 2 * h + 1 <= idx < 3 * h + 1 and bool(self._mask[0, b, idx])

    @staticmethod
    def _message_byte(agent: str, m) -> int:
        m = np.asarray(m).astype(bool)
        assert m.shape == (MESSAGE_LENGTH,), f"{agent} attempting to send message {m} that is not in the message space"  # SimulationController.py:797
        return int(sum(int(x) << i for i, x in enumerate(m)))

    def action_index(self, agent: str, action) -> int:
        """The BlueFixedActionWrapper index of an action given as an index, as a descriptor of `actions()`, or as any object that
        looks like one of the reference's blue Action objects (class name + `hostname` or `from_subnet` / `to_subnet` attributes:
        Analyse / Remove / Restore / DeployDecoy / Monitor / Sleep / BlockTrafficZone / AllowTrafficZone)."""
        if isinstance(action, (int, np.integer)):
            return int(action)
        name = action.name if isinstance(action, ActionDescriptor) else type(action).__name__
        want = {k: getattr(action, k) for k in ("hostname", "from_subnet", "to_subnet") if getattr(action, k, None) is not None}
        b = int(agent.rsplit("_", 1)[1])
        if name == "Sleep":
            return SLEEP_INDEX[b]
        if name == "Monitor":
            return 48 if b == 4 else 16
        for i, d in enumerate(self._labels().actions(agent)):
            if want and d.name == name and all(d.params.get(k) == v for k, v in want.items()):
                return i
        raise ValueError(f"{agent}: {name} {want} is not one of the agent's actions (BlueFixedActionWrapper.actions); a host that does not exist in this "
                         f"episode has no action, and only blue actions can be submitted")

    # env.py:95
    def parallel_step(self, actions: dict | None = None, messages: dict | None = None, skip_valid_action_check: bool = False):
        """`actions`: {blue agent: BlueFixedActionWrapper index, or a descriptor / reference-style Action object that `action_index`
        can map to one} (the engine executes indices; agents left out Sleep, as the reference's SleepAgent default does).  Returns (observations, rewards, dones, {}) with the reference's
        structure: dict observations (see get_observation), the team reward dict per agent, one done flag per agent."""
        self._single("CybORG.parallel_step")
        if self._tracker is None:
            if self._ticks:
                raise RuntimeError("parallel_step returns dict observations, which follow the episode from its first tick: call "
                                   "enable_dict_observations() (or parallel_step) before the episode's first tick")
            self.enable_dict_observations()
        idx = {a: self.action_index(a, v) for a, v in (actions or {}).items()}
        self._tick(idx, {a: self._message_byte(a, m) for a, m in (messages or {}).items()})
        agents = list(dict.fromkeys(list(idx) + self.active_agents))
        return ({a: self.get_observation(a) for a in agents}, {a: dict(self._rewards["Blue"]) for a in agents},
                {a: self._done for a in agents}, {})

    # env.py:125 (CC1 / CC2 compatibility call)
    def step(self, agent: str | None = None, action=None, messages: dict | None = None, skip_valid_action_check: bool = False) -> Results:
        obs, rew, done, _ = self.parallel_step({} if agent is None or action is None else {agent: action}, messages)
        who = agent if agent is not None else self._blue[0]
        return Results(observation=obs[who], done=done[who], reward=round(sum(rew[who].values()), 1), action_space=self.get_action_space(who), action=action)

    # env.py:218
    def reset(self, agent: str | None = None, seed: int | None = None):
        if seed is not None:
            self.seed, self._episodes_started = int(seed), 0
        self.engine.set_episode_base(self.seed, self._episodes_started)
        self._episodes_started += self.engine.num_envs
        self._obs, self._mask = self.engine.reset()
        self._last_actions = None
        self._done = False
        self._ticks = 0
        self._rewards["Blue"] = {"BlueRewardMachine": 0.0, "action_cost": 0.0}
        if self._want_dict:
            self.enable_dict_observations()
        if agent is None:
            return None
        return Results(observation=self.get_observation(agent), action_space=self.get_action_space(agent))

    def set_seed(self, seed: int):
        self.seed, self._episodes_started = int(seed), 0

    def enable_dict_observations(self):
        """From the next reset on (immediately when the episode has not started), follow the episode tick by tick so that
        get_observation can rebuild the reference's dict observations (one state export per tick: a debugging / logging aid for
        ONE episode, not a throughput path).  parallel_step and BlueFixedActionWrapper switch it on."""
        self._single("dict observations")
        self._want_dict = True
        if self._ticks == 0:
            from .observations import BlueObservationTracker
            self._tracker = BlueObservationTracker(self)

    # ---- read-only views ---------------------------------------------------------------------------------------------
    @property
    def agents(self) -> list[str]:
        """env.py:394: the external-facing agents.  The green and red agents are scripted inside the kernel."""
        return list(self._blue)

    @property
    def active_agents(self) -> list[str]:
        return list(self._blue)

    def get_agent_ids(self) -> list[str]:
        return list(self._blue)

    def get_message_space(self, agent: str):
        return MultiBinary(MESSAGE_LENGTH)

    def get_rewards(self) -> dict:
        """env.py:348: the team rewards of the last tick, {'Blue': {'BlueRewardMachine': ., 'action_cost': .}, 'Red': ..., 'Green': ...}
        (EnterpriseScenarioGenerator._generate_team_calcs :819-830: only the blue team has a reward calculator)."""
        return {k: dict(v) for k, v in self._rewards.items()}

    def get_attr(self, attribute: str) -> Any:
        """env.py:374."""
        return getattr(self, attribute) if hasattr(self, attribute) else None

    def _hostnames(self, env_index: int = 0) -> list[str]:
        """Existing hosts of the episode in state.hosts order (creation order), from the reset-time masks."""
        st = self.get_agent_state("True", env_index)
        order = []
        for sn in SUBNETS_CREATION[:8]:
            order += [h for h in subnet_hosts(sn) if h in st["hosts"]]
        return order + ["root_internet_host_0"]

    def get_ip_map(self) -> dict:
        """env.py:327: hostname -> IPv4Address (synthetic numbering, see host_ip)."""
        return {h: host_ip(h) for h in self._hostnames()}

    def get_cidr_map(self) -> dict:
        """env.py:337: subnet name -> IPv4Network (synthetic numbering)."""
        return {sn: subnet_cidr(sn) for sn in SUBNETS_CREATION}

    def get_last_action(self, agent: str):
        """The action the agent submitted last tick, as a one-element list like env.py:300 returns: an int (the
        BlueFixedActionWrapper index) that prints as the action's label, e.g. `Restore server_host_0`, which is what the
        reference's Action objects print in the evaluation logs (evaluation.py:111-117)."""
        if self._last_actions is None:
            return None
        b = int(agent.rsplit("_", 1)[1])
        idx = int(self._last_actions[0, b])
        labels = self._label_source(agent) if self._label_source is not None else self._labels().action_labels(agent)
        return [LastAction(idx, labels[idx] if idx < len(labels) else str(idx))]

    def _labels(self):
        if self._helper is None:
            self._helper = BlueFixedActionWrapper.__new__(BlueFixedActionWrapper)
            self._helper.env, self._helper._pad_spaces, self._helper.env_index = self, False, 0
        self._helper._refresh_masks()
        return self._helper

    def get_action_space(self, agent: str) -> dict:
        """env.py:266 for a blue agent: the dict-of-dicts action space (`action` classes, `subnet`, `ip_address`, `hostname`,
        `session`, `agent` -> known flag; AgentInterface.py:105-117, ActionSpace.py:105).  A blue agent knows its own subnets'
        hosts from the start (its sessions' observations), nothing else, and that never changes during an episode."""
        b = int(agent.rsplit("_", 1)[1])
        mine = set()
        st = self.get_agent_state("True")
        for sn in BLUE_SUBNETS[b]:
            mine |= {h for h in subnet_hosts(sn) if h in st["hosts"]}
        hosts = self._hostnames()
        names = ["Sleep", "Monitor", "Analyse", "Remove", "DeployDecoy", "Restore", "BlockTrafficZone", "AllowTrafficZone"]   # EnterpriseScenarioGenerator.py:642
        return {"action": {n: True for n in names}, "allowed_subnets": list(BLUE_SUBNETS[b]),
                "subnet": {subnet_cidr(sn): sn in BLUE_SUBNETS[b] for sn in SUBNETS_CREATION},
                "ip_address": {host_ip(h): h in mine for h in hosts}, "hostname": {h: h in mine for h in hosts},
                "session": {0: True}, "agent": {agent: True}, "username": {}, "password": {}, "process": {}, "port": {}, "target_session": {}}

    def get_observation(self, agent: str, env_index: int = 0) -> dict:
        """env.py:245: the agent's last observation as the reference's nested dict, rebuilt for ONE episode from what the engine
        tracks (see observations.blue_observation): `success`, `action`, `message`, and one entry per host of the agent's
        subnets on which this tick's Monitor found events."""
        from .observations import blue_observation
        return blue_observation(self, agent, env_index)

    def get_agent_state(self, agent: str = "True", env_index: int = 0) -> dict:
        """env.py:265 `get_agent_state('True')`: the true state of one episode as a nested dict (see state.true_state; keyed by
        hostnames, listing processes, services, sessions and events); for a blue agent name: its dict observation."""
        if agent != "True":
            return self.get_observation(agent, env_index)
        from .state import true_state
        return true_state(self.engine.export_canon(env_index))

    @property
    def unwrapped(self):
        return self


class BlueFixedActionWrapper:
    """BlueFixedActionWrapper.py:44 surface: fixed discrete action indices with labels and masks."""

    def __init__(self, env: CybORG, pad_spaces: bool = False, env_index: int = 0, *args, **kwargs):
        env._single(type(self).__name__)
        assert env_index == 0
        if type(self) is BlueFixedActionWrapper:        # this class returns the environment's dict observations (BlueFixedActionWrapper.py:101-185)
            env.enable_dict_observations()
        self.env, self._pad_spaces, self.env_index = env, pad_spaces, env_index
        self.agents = self.possible_agents = [f"blue_agent_{i}" for i in range(5)]
        self._refresh_masks()
        env._label_source = self._cached_labels

    # -- metadata (BlueFixedActionWrapper._create_hardcoded_metadata :200-231) ------------------------------------------
    def subnets(self, agent_name: str) -> list[str]:
        return sorted(BLUE_SUBNETS[self._idx(agent_name)])

    def hosts(self, agent_name: str) -> list[str]:
        out = []
        for sn in BLUE_SUBNETS[self._idx(agent_name)]:
            out += subnet_hosts(sn)
        return sorted(out)

    @staticmethod
    def _idx(agent_name: str) -> int:
        return int(agent_name.rsplit("_", 1)[1])

    def _refresh_masks(self):
        self._mask = self.env._mask[self.env_index].cpu().numpy().astype(bool)
        self._labels = {}
        self._actions = {}

    def _cached_labels(self, agent_name: str) -> list[str]:
        """Labels of the current episode (they only change with the masks, i.e. at reset)."""
        if agent_name not in self._labels:
            self._labels[agent_name] = self.action_labels(agent_name)
        return self._labels[agent_name]

    def action_mask(self, agent_name: str) -> list[bool]:
        b = self._idx(agent_name)
        n = 242 if self._pad_spaces else N_ACTIONS[b]
        return self._mask[b, :n].tolist()

    def _enumerate(self, agent_name: str):
        """(label, descriptor) of every index in the order of BlueFixedActionWrapper._populate_action_space :233-309: commands
        sorted by class name, hosts sorted by name (routers excluded), absent hosts as masked-out Sleep."""
        b = self._idx(agent_name)
        hosts = [h for h in self.hosts(agent_name) if "router" not in h]
        mask = self._mask[b]
        out = []
        params = {"session": 0, "agent": agent_name}

        def host_block(cmd):
            for h in hosts:
                out.append((f"{cmd} {h}", ActionDescriptor(cmd, hostname=h, **params)) if mask[len(out)] else (f"[Invalid] {cmd} {h}", ActionDescriptor("Sleep")))

        def zone_block(cmd):
            for dst in self.subnets(agent_name):
                for src in SUBNETS_SORTED:
                    if src != dst:
                        out.append((f"{cmd} {dst} ({subnet_cidr(dst)}) <- {src} ({subnet_cidr(src)})", ActionDescriptor(cmd, from_subnet=src, to_subnet=dst, **params)))

        host_block("Analyse")
        out.append(("Monitor", ActionDescriptor("Monitor", **params)))
        host_block("Remove")
        host_block("Restore")
        out.append(("Sleep", ActionDescriptor("Sleep")))
        zone_block("AllowTrafficZone")
        zone_block("BlockTrafficZone")
        host_block("DeployDecoy")
        if self._pad_spaces:
            out += [("[Padding] Sleep", ActionDescriptor("Sleep"))] * (242 - len(out))
        return out

    def action_labels(self, agent_name: str) -> list[str]:
        """BlueFixedActionWrapper.action_labels :355 (the CIDRs in the traffic-control labels are the synthetic ones)."""
        return [lab for lab, _ in self._enumerate(agent_name)]

    def actions(self, agent_name: str) -> list[ActionDescriptor]:
        """BlueFixedActionWrapper.actions :360: one descriptor per index (see ActionDescriptor)."""
        if agent_name not in self._actions:
            self._actions[agent_name] = [a for _, a in self._enumerate(agent_name)]
        return self._actions[agent_name]

    @property
    def is_padded(self) -> bool:
        return self._pad_spaces

    @functools.lru_cache(maxsize=None)
    def action_space(self, agent_name: str):
        return Discrete(242 if self._pad_spaces else N_ACTIONS[self._idx(agent_name)])

    def action_spaces(self):
        return {a: self.action_space(a) for a in self.possible_agents}

    def get_action_space(self, agent: str):
        return {"actions": self.actions(agent), "labels": self.action_labels(agent), "mask": self.action_mask(agent)}

    # the wrappers forward everything else to the environment, as BaseWrapper.get_attr does (BaseWrapper.py:8)
    def get_attr(self, attribute: str):
        return getattr(self, attribute) if hasattr(self, attribute) else self.env.get_attr(attribute)

    def get_observation(self, agent: str):
        return self.env.get_observation(agent)

    def get_last_action(self, agent: str):
        return self.env.get_last_action(agent)

    def get_rewards(self):
        return self.env.get_rewards()

    def get_ip_map(self):
        return self.env.get_ip_map()

    def get_message_space(self, agent: str):
        return self.env.get_message_space(agent)

    @property
    def unwrapped(self):
        return self.env

    # BlueFixedActionWrapper.reset :76 / step :101 (dict observations of the environment)
    def reset(self, *args, **kwargs):
        self.env.reset(*args, **kwargs)
        self.agents = list(self.possible_agents)
        self._refresh_masks()
        return {a: self.env.get_observation(a) for a in self.possible_agents}, self._info()

    def _info(self):
        return {a: {"action_mask": self.action_mask(a)} for a in self.possible_agents}

    def _submit(self, actions, messages):
        idx = {}
        for a, v in (actions or {}).items():
            if not isinstance(v, (int, np.integer)):
                raise TypeError(f"{a}: the engine takes BlueFixedActionWrapper indices, got {type(v).__name__}")
            b = self._idx(a)
            if not 0 <= int(v) < (242 if self._pad_spaces else N_ACTIONS[b]):
                raise IndexError(f"{a}: action index {v} out of range")
            idx[a] = int(v)
        obs_row, r, d = self.env._tick(idx, {a: CybORG._message_byte(a, m) for a, m in (messages or {}).items()})
        rewards = {a: r for a in self.possible_agents}
        terminated = {a: d for a in self.possible_agents}
        truncated = {a: d for a in self.possible_agents}
        self.agents = [] if d else list(self.possible_agents)          # BlueFixedActionWrapper.py:165-167
        return obs_row, rewards, terminated, truncated

    def step(self, actions: dict[str, int] | None = None, messages: dict[str, Any] | None = None, **kwargs):
        _, rewards, terminated, truncated = self._submit(actions, messages)
        return {a: self.env.get_observation(a) for a in self.possible_agents}, rewards, terminated, truncated, self._info()


class BlueFlatWrapper(BlueFixedActionWrapper):
    """BlueFlatWrapper.py:40: fixed-length vector observations (92 / 210 entries, int64 like the reference)."""

    # -- spaces (BlueFlatWrapper._get_init_obs_spaces :128-170) ---------------------------------------------------------
    @functools.lru_cache(maxsize=None)
    def observation_space(self, agent_name: str):
        middle = list(itertools.chain(NUM_SUBNETS * [2], NUM_SUBNETS * [2], NUM_SUBNETS * [2], MAX_HOSTS * [2], MAX_HOSTS * [2]))
        tail = (NUM_MESSAGES * MESSAGE_LENGTH) * [2]
        long = self._pad_spaces or agent_name == "blue_agent_4"
        return MultiDiscrete([3] + (3 if long else 1) * middle + tail)

    def observation_spaces(self):
        return {a: self.observation_space(a) for a in self.possible_agents}

    def _obs_dict(self, obs_row: np.ndarray) -> dict[str, np.ndarray]:
        out = {}
        for b, a in enumerate(self.possible_agents):
            v = obs_row[b, :OBS_SIZES[b]].astype(np.int64)
            if self._pad_spaces and len(v) < 210:
                v = np.pad(v, (0, 210 - len(v)))       # BlueFlatWrapper.py:251-254
            out[a] = v
        return out

    # BlueFlatWrapper.reset :61
    def reset(self, *args, **kwargs):
        self.env.reset(*args, **kwargs)
        self.agents = list(self.possible_agents)
        self._refresh_masks()
        return self._obs_dict(self.env._obs[self.env_index].cpu().numpy()), self._info()

    # BlueFlatWrapper.step :80 / BlueFixedActionWrapper.step :101
    def step(self, actions: dict[str, int] | None = None, messages: dict[str, Any] | None = None, **kwargs):
        obs_row, rewards, terminated, truncated = self._submit(actions, messages)
        return self._obs_dict(obs_row), rewards, terminated, truncated, self._info()


class BlueEnterpriseWrapper(BlueFlatWrapper):
    """BlueEnterpriseWrapper.py:13: also accepts {"actions": ..., "messages": ...} and reset(agent, seed)."""

    def step(self, actions=None, messages=None, **kwargs):
        action_dict = actions if actions is not None else {}
        if "actions" in action_dict:
            messages = action_dict.get("messages", messages)
            return super().step(action_dict["actions"], messages=messages)
        return super().step(action_dict, messages=messages)

    def reset(self, agent=None, seed=None, *args, **kwargs):
        return super().reset(agent=agent, seed=seed)

    @property
    def long_observation_space(self):
        return self.observation_space("blue_agent_4")

    @property
    def short_observation_space(self):
        return self.observation_space("blue_agent_0")


class EnterpriseMAE(BlueEnterpriseWrapper):
    """EnterpriseMAE.py:10 (RLlib MultiAgentEnv shape): adds the "__all__" keys."""

    def step(self, action_dict=None, messages=None):
        obs, rew, terminated, truncated, info = BlueFlatWrapper.step(
            self, **self._split(action_dict, messages))
        terminated["__all__"] = False
        truncated["__all__"] = bool(next(iter(truncated.values())))       # determine_done, EnterpriseMAE.py:70-71
        return obs, rew, terminated, truncated, info

    @staticmethod
    def _split(action_dict, messages):
        action_dict = action_dict if action_dict is not None else {}
        if "actions" in action_dict:
            return dict(actions=action_dict["actions"], messages=action_dict.get("messages", messages))
        return dict(actions=action_dict, messages=messages)
