"""The evaluation protocol of the reference (CybORG/Evaluation/evaluation.py:56-134) on the batched engine.

`run_evaluation` is the reference loop, one episode at a time, through the mirrored wrapper API (a submission object
with NAME / TEAM / TECHNIQUE / AGENTS / wrap works unchanged).  `run_evaluation_batched` is the same protocol with all
episodes stepped concurrently and the policy evaluated on device; across ranks the per-episode totals are exchanged
with ONE collective at the end (the only collective on the path).
"""
from __future__ import annotations

import time
from statistics import mean, stdev

import numpy as np
import torch

from .engine import BatchedCC4, N_ACTIONS, OBS_SIZES
from .wrappers import CybORG, EnterpriseGreenAgent, EnterpriseScenarioGenerator, FiniteStateRedAgent, SleepAgent

EPISODE_LENGTH = 500


def run_evaluation(submission, log_path: str | None = None, max_eps: int = 100, write_to_file: bool = False, seed=None, device: int = 0):
    """evaluation.py:56-134 with the same bookkeeping: per-episode sum of the mean blue reward, the reward of the tick
    on which `done` fires is not counted (evaluation.py:104-110)."""
    sg = EnterpriseScenarioGenerator(blue_agent_class=SleepAgent, green_agent_class=EnterpriseGreenAgent,
                                     red_agent_class=FiniteStateRedAgent, steps=EPISODE_LENGTH)
    cyborg = CybORG(sg, "sim", seed=seed, device=device)
    wrapped_cyborg = submission.wrap(cyborg)
    start = time.time()
    total_reward, actions_log = [], []
    for _ in range(max_eps):
        observations, _ = wrapped_cyborg.reset()
        r, a = [], []
        for _ in range(EPISODE_LENGTH):
            actions = {name: agent.get_action(observations[name], wrapped_cyborg.action_space(name))
                       for name, agent in submission.AGENTS.items() if name in wrapped_cyborg.agents}
            observations, rew, term, trunc, info = wrapped_cyborg.step(actions)
            done = {agent: term.get(agent, False) or trunc.get(agent, False) for agent in wrapped_cyborg.possible_agents}
            if all(done.values()):
                break
            r.append(mean(rew.values()))
            if write_to_file:
                a.append({name: cyborg.get_last_action(name) for name in wrapped_cyborg.agents})
        total_reward.append(sum(r))
        actions_log.append(a)
    out = {"mean": mean(total_reward), "stdev": stdev(total_reward) if len(total_reward) > 1 else 0.0,
           "episodes": len(total_reward), "seconds": time.time() - start, "rewards": total_reward}
    if write_to_file and log_path:
        import json
        import os
        os.makedirs(log_path, exist_ok=True)
        with open(os.path.join(log_path, "summary.json"), "w") as f:
            json.dump({k: v for k, v in out.items() if k != "rewards"}, f)
        with open(os.path.join(log_path, "scores.txt"), "w") as f:
            f.write(f"reward_mean: {out['mean']}\nreward_stdev: {out['stdev']}\n")
    return out


class MLPPolicy(torch.nn.Module):
    """BASELINE config 5: one random-init MLP per blue agent (obs -> 256 -> 256 -> n_actions), masked argmax."""

    def __init__(self, hidden: int = 256, seed: int = 0):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.nets = torch.nn.ModuleList()
        for b in range(5):
            net = torch.nn.Sequential(torch.nn.Linear(OBS_SIZES[b], hidden), torch.nn.ReLU(), torch.nn.Linear(hidden, hidden),
                                      torch.nn.ReLU(), torch.nn.Linear(hidden, N_ACTIONS[b]))
            for p in net.parameters():
                with torch.no_grad():
                    p.copy_((torch.rand(p.shape, generator=g) * 2 - 1) / (p.shape[-1] ** 0.5))
            self.nets.append(net)

    @torch.no_grad()
    def forward(self, obs: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
        acts = []
        for b, net in enumerate(self.nets):
            logits = net(obs[:, b, :OBS_SIZES[b]].float())
            logits = logits.masked_fill(~mask[:, b, :N_ACTIONS[b]], float("-inf"))
            acts.append(logits.argmax(-1))
        return torch.stack(acts, 1).to(torch.int32).contiguous()


class RandomBluePolicy:
    """The reference's cc4BlueRandomAgent (CybORG/Agents/SimpleAgents/RandomAgent.py:66-69) as a device-side batched
    policy: every blue agent draws uniformly among its VALID actions (mask True), BlockTrafficZone / AllowTrafficZone
    excluded as in the reference (`excluded_actions`).  `policy(obs, mask) -> int32 [N, 5]`."""

    def __init__(self, seed: int = 0, exclude_traffic_control: bool = True):
        self.seed, self.exclude, self._g = seed, exclude_traffic_control, None

    @torch.no_grad()
    def __call__(self, obs: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
        if self._g is None:
            self._g = torch.Generator(device=mask.device).manual_seed(self.seed)
        m = mask.clone()
        if self.exclude:     # index layout of BlueFixedActionWrapper._populate_action_space :233-309: Allow + Block follow Sleep
            m[:, :4, 50:66] = False
            m[:, 4, 146:194] = False
        key = torch.rand(m.shape, generator=self._g, device=m.device).masked_fill(~m, -1.0)
        return key.argmax(-1).to(torch.int32).contiguous()


def shard(n_total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous episode range [lo, hi) owned by `rank` (episodes are keyed by global id, so totals do not depend on world)."""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def aggregate_scores(local_totals: torch.Tensor, n_total: int, rank: int, world: int) -> dict:
    """All ranks' per-episode totals -> mean / stdev exactly as statistics.mean / stdev over the full list
    (evaluation.py:133-134).  The all_gather below is the only collective of the whole path."""
    import torch.distributed as dist
    if world > 1:
        sizes = [shard(n_total, r, world)[1] - shard(n_total, r, world)[0] for r in range(world)]
        pad = torch.zeros(max(sizes), dtype=torch.float64, device=local_totals.device)
        pad[:local_totals.numel()] = local_totals.double()
        bufs = [torch.zeros_like(pad) for _ in range(world)]
        dist.all_gather(bufs, pad)
        totals = torch.cat([b[:n] for b, n in zip(bufs, sizes)])
    else:
        totals = local_totals.double()
    vals = totals.cpu().tolist()
    return {"mean": mean(vals), "stdev": stdev(vals) if len(vals) > 1 else 0.0, "episodes": len(vals), "rewards": vals}


def run_evaluation_batched(policy, max_eps: int = 100, seed: int = 0, device: int = 0, rank: int = 0, world: int = 1,
                           steps: int = EPISODE_LENGTH) -> dict:
    """The protocol with every episode of this rank's shard stepped concurrently; `policy(obs, mask) -> int32 [N,5]`
    runs on device (no host round trip per tick)."""
    lo, hi = shard(max_eps, rank, world)
    n = hi - lo
    dev = torch.device("cuda", device)
    totals = torch.zeros(max(n, 0), dtype=torch.float64, device=dev)
    t0 = time.time()
    if n > 0:
        env = BatchedCC4(n, device=device, seed=seed, steps=steps, first_episode=lo)
        obs, mask = env.reset()
        for _ in range(steps):
            acts = policy(obs, mask)
            obs, rew, done = env.step(acts)
            if bool(done[0]):          # every episode has the same length; the done tick's reward is dropped
                break
            totals += rew.double()
    out = aggregate_scores(totals, max_eps, rank, world)
    out["seconds"] = time.time() - t0
    return out
