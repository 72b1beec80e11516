"""Host-side mirror of the reference's environment / wrapper API for the hot path, backed by the CUDA engine.

Same names, argument meaning and return structure as the reference classes, so code written against
`CybORG(...)` + `BlueFlatWrapper` / `BlueEnterpriseWrapper` / `EnterpriseMAE` (reference: CybORG/env.py:53,
CybORG/Agents/Wrappers/BlueFixedActionWrapper.py:44, BlueFlatWrapper.py:40, BlueEnterpriseWrapper.py:13,
EnterpriseMAE.py:10) — in particular the loop of CybORG/Evaluation/evaluation.py:56-134 — runs unchanged on it.
Each wrapper object presents ONE episode of a `BatchedCC4` batch (index `env_index`); the batched API itself is
`cage_challenge_4_b200.BatchedCC4`.

What is NOT mirrored (out of the hot path): dict observations from `CybORG.get_observation`,
`Action` objects (actions are the wrapper's integer indices), IP addresses / CIDRs in labels, user-supplied red / green
agent classes (the engine scripts FiniteStateRedAgent and EnterpriseGreenAgent; anything else raises).
Seeds select Philox streams, not numpy PCG64 streams, so a seeded trajectory is reproducible but differs from the
reference's trajectory for the same integer (parity is established through RNG-tape replay, tests/).
"""
from __future__ import annotations

import functools
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


class EnterpriseScenarioGenerator:
    """Same constructor as Simulator/Scenarios/EnterpriseScenarioGenerator.py:95-101."""
    MIN_USER_HOSTS, MAX_USER_HOSTS, MIN_SERVER_HOSTS, MAX_SERVER_HOSTS = 3, 10, 1, 6
    MAX_BANDWIDTH, MESSAGE_LENGTH = 100, 8

    def __init__(self, blue_agent_class=None, red_agent_class=None, green_agent_class=None, steps: int = 100):
        red_name = "FiniteStateRedAgent" if red_agent_class is None else getattr(red_agent_class, "__name__", "")
        if red_name not in ("FiniteStateRedAgent", "DiscoveryFSRed"):
            raise NotImplementedError("the engine scripts FiniteStateRedAgent and DiscoveryFSRed (SURVEY.md section 8f lists the other red agents as next)")
        self.red_agent = red_name
        if green_agent_class not in (None, EnterpriseGreenAgent) and getattr(green_agent_class, "__name__", "") != "EnterpriseGreenAgent":
            raise NotImplementedError("the engine scripts EnterpriseGreenAgent only")
        self.blue_agent_class, self.red_agent_class, self.green_agent_class, self.steps = blue_agent_class, red_agent_class, green_agent_class, steps


class LastAction(int):
    """An action index that prints as the action's label."""

    def __new__(cls, index: int, label: str):
        self = super().__new__(cls, index)
        self.label = label
        return self

    def __repr__(self):
        return self.label

    __str__ = __repr__


class CybORG:
    """Facade with the constructor of CybORG/env.py:53-77.  `agents` is accepted and ignored exactly as the evaluation
    script uses it (it passes "sim", evaluation.py:70).  Extra keywords choose the batch: num_envs episodes on `device`."""

    def __init__(self, scenario_generator: EnterpriseScenarioGenerator, agents=None, seed: int | None = None, *,
                 num_envs: int = 1, device: int = 0):
        assert isinstance(scenario_generator, EnterpriseScenarioGenerator), "scenario_generator must be an EnterpriseScenarioGenerator"
        self.scenario_generator = scenario_generator
        self.seed = int(np.random.SeedSequence().entropy % (2**63)) if seed is None else int(seed)
        self.engine = BatchedCC4(num_envs, device=device, seed=self.seed, steps=scenario_generator.steps, red_agent=scenario_generator.red_agent)
        self._episodes_started = 0
        self._last_actions = None
        self._label_source = None       # set by the wrapper: agent name -> list of action labels
        self.agents = [f"blue_agent_{i}" for i in range(5)]
        self.reset()

    # env.py:218
    def reset(self, agent: str | None = None, seed: int | None = None):
        if seed is not None:
            self.seed, self._episodes_started = int(seed), 0
        self.engine.set_episode_base(self.seed, self._episodes_started)
        self._episodes_started += self.engine.num_envs
        self._obs, self._mask = self.engine.reset()
        self._last_actions = None
        return None

    def set_seed(self, seed: int):
        self.seed, self._episodes_started = int(seed), 0

    @property
    def active_agents(self):
        return list(self.agents)

    def get_message_space(self, agent: str):
        return MultiBinary(MESSAGE_LENGTH)

    def get_last_action(self, agent: str):
        """The action the agent submitted last tick, as a one-element list like env.py:300 returns: an int (the
        BlueFixedActionWrapper index) that prints as the action's label, e.g. `Restore server_host_0`, which is what the
        reference's Action objects print in the evaluation logs (evaluation.py:111-117)."""
        if self._last_actions is None:
            return None
        b = int(agent.rsplit("_", 1)[1])
        idx = int(self._last_actions[0, b])
        labels = self._label_source(agent) if self._label_source is not None else None
        return [LastAction(idx, labels[idx] if labels is not None and idx < len(labels) else str(idx))]

    def get_agent_state(self, agent: str = "True", env_index: int = 0) -> dict:
        """env.py:265 `get_agent_state('True')`: the true state of one episode as a nested dict (see state.true_state; the
        engine has no IP addresses or files, so the dict is keyed by hostnames and lists processes, services, sessions and
        events).  Only the true state is available: per-agent dict observations are not reconstructed."""
        if agent != "True":
            raise NotImplementedError("only get_agent_state('True') is available; blue agents observe through the BlueFlatWrapper vectors")
        from .state import true_state
        return true_state(self.engine.export_canon(env_index))

    @property
    def unwrapped(self):
        return self


class BlueFixedActionWrapper:
    """BlueFixedActionWrapper.py:44 surface: fixed discrete action indices with labels and masks."""

    def __init__(self, env: CybORG, pad_spaces: bool = False, env_index: int = 0, *args, **kwargs):
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
            out.append(f"{sn}_router")
            out += [f"{sn}_user_host_{i}" for i in range(MAX_USER_HOSTS)]
            out += [f"{sn}_server_host_{i}" for i in range(MAX_SERVER_HOSTS)]
        return sorted(out)

    @staticmethod
    def _idx(agent_name: str) -> int:
        return int(agent_name.rsplit("_", 1)[1])

    def _refresh_masks(self):
        self._mask = self.env._mask[self.env_index].cpu().numpy().astype(bool)
        self._labels = {}

    def _cached_labels(self, agent_name: str) -> list[str]:
        """Labels of the current episode (they only change with the masks, i.e. at reset)."""
        if agent_name not in self._labels:
            self._labels[agent_name] = self.action_labels(agent_name)
        return self._labels[agent_name]

    def action_mask(self, agent_name: str) -> list[bool]:
        b = self._idx(agent_name)
        n = 242 if self._pad_spaces else N_ACTIONS[b]
        return self._mask[b, :n].tolist()

    def action_labels(self, agent_name: str) -> list[str]:
        """BlueFixedActionWrapper._populate_action_space :233-309 (labels carry no CIDRs: addresses are not modelled)."""
        b = self._idx(agent_name)
        hosts = [h for h in self.hosts(agent_name) if "router" not in h]
        mask = self._mask[b]
        labels, k = [], 0

        def host_block(cmd):
            nonlocal k
            for h in hosts:
                labels.append(f"{cmd} {h}" if mask[k] else f"[Invalid] {cmd} {h}")
                k += 1

        def zone_block(cmd):
            nonlocal k
            for dst in self.subnets(agent_name):
                for src in SUBNETS_SORTED:
                    if src != dst:
                        labels.append(f"{cmd} {dst} <- {src}")
                        k += 1

        host_block("Analyse")
        labels.append("Monitor"); k += 1
        host_block("Remove")
        host_block("Restore")
        labels.append("Sleep"); k += 1
        zone_block("AllowTrafficZone")
        zone_block("BlockTrafficZone")
        host_block("DeployDecoy")
        if self._pad_spaces:
            labels += ["[Padding] Sleep"] * (242 - len(labels))
        return labels

    @property
    def is_padded(self) -> bool:
        return self._pad_spaces

    @functools.lru_cache(maxsize=None)
    def action_space(self, agent_name: str):
        return Discrete(242 if self._pad_spaces else N_ACTIONS[self._idx(agent_name)])

    def action_spaces(self):
        return {a: self.action_space(a) for a in self.possible_agents}

    def get_action_space(self, agent: str):
        return {"labels": self.action_labels(agent), "mask": self.action_mask(agent)}

    @property
    def unwrapped(self):
        return self.env


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
                v = np.pad(v, (0, 210 - len(v)))
            out[a] = v
        return out

    def _info(self):
        return {a: {"action_mask": self.action_mask(a)} for a in self.possible_agents}

    # BlueFlatWrapper.reset :61
    def reset(self, *args, **kwargs):
        self.env.reset(*args, **kwargs)
        self.agents = list(self.possible_agents)
        self._refresh_masks()
        return self._obs_dict(self.env._obs[self.env_index].cpu().numpy()), self._info()

    # BlueFlatWrapper.step :80 / BlueFixedActionWrapper.step :101
    def step(self, actions: dict[str, int] | None = None, messages: dict[str, Any] | None = None, **kwargs):
        eng = self.env.engine
        acts = torch.tensor([SLEEP_INDEX] * eng.num_envs, dtype=torch.int32)
        for a, v in (actions or {}).items():
            if not isinstance(v, (int, np.integer)):
                raise TypeError(f"{a}: the engine takes BlueFixedActionWrapper indices, got {type(v).__name__}")
            b = self._idx(a)
            if not 0 <= int(v) < (242 if self._pad_spaces else N_ACTIONS[b]):
                raise IndexError(f"{a}: action index {v} out of range")
            acts[self.env_index, b] = int(v)
        msgs = torch.zeros((eng.num_envs, 5), dtype=torch.uint8)
        for a, m in (messages or {}).items():
            m = np.asarray(m).astype(bool)
            assert m.shape == (MESSAGE_LENGTH,), f"{a} attempting to send message {m} that is not in the message space"  # SimulationController.py:797
            msgs[self.env_index, self._idx(a)] = int(sum(int(x) << i for i, x in enumerate(m)))
        obs, rew, done = eng.step(acts.to(eng.device), msgs.to(eng.device))
        self.env._last_actions = acts[self.env_index:self.env_index + 1].numpy()
        obs_row = obs[self.env_index].cpu().numpy()
        r, d = float(rew[self.env_index]), bool(done[self.env_index])
        rewards = {a: r for a in self.possible_agents}
        terminated = {a: d for a in self.possible_agents}
        truncated = {a: d for a in self.possible_agents}
        self.agents = [] if d else list(self.possible_agents)          # BlueFixedActionWrapper.py:165-167
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
