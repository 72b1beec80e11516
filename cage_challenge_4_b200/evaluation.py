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
    """evaluation.py:56-197 with the same bookkeeping and the same outputs: per-episode sum of the mean blue reward, the reward
    of the tick on which `done` fires is not counted (evaluation.py:104-110); with write_to_file the five log files of
    evaluation.py:141-196 (summary.txt, full.txt, actions.txt, summary.json, scores.txt) are written to log_path."""
    from datetime import datetime
    sg = EnterpriseScenarioGenerator(blue_agent_class=SleepAgent, green_agent_class=EnterpriseGreenAgent,
                                     red_agent_class=FiniteStateRedAgent, steps=EPISODE_LENGTH)
    cyborg = CybORG(sg, "sim", seed=seed, device=device)
    wrapped_cyborg = submission.wrap(cyborg)
    version_header = f"cage_challenge_4_b200 (CybORG v4.0 semantics), Scenario4"
    author_header = f"Author: {getattr(submission, 'NAME', '?')}, Team: {getattr(submission, 'TEAM', '?')}, Technique: {getattr(submission, 'TECHNIQUE', '?')}"
    start = datetime.now()
    total_reward, actions_log, obs_log = [], [], []
    for _ in range(max_eps):
        observations, _ = wrapped_cyborg.reset()
        r, a, o = [], [], []
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
                o.append({name: observations[name] for name in observations.keys()})
        total_reward.append(sum(r))
        if write_to_file:
            actions_log.append(a)
            obs_log.append(o)
    end = datetime.now()
    out = {"mean": mean(total_reward), "stdev": stdev(total_reward) if len(total_reward) > 1 else 0.0,
           "episodes": len(total_reward), "seconds": (end - start).total_seconds(), "rewards": total_reward}
    if write_to_file and log_path:
        import json
        import os
        os.makedirs(log_path, exist_ok=True)
        reward_string = f"Average reward is: {out['mean']} with a standard deviation of {out['stdev']}"
        with open(os.path.join(log_path, "summary.txt"), "w") as f:
            f.write(version_header + "\n" + author_header + "\n" + reward_string + "\n" + f"Using agents {submission.AGENTS}")
        with open(os.path.join(log_path, "full.txt"), "w") as f:
            f.write(version_header + "\n" + author_header + "\n" + reward_string + "\n")
            for act, obs, sum_rew in zip(actions_log, obs_log, total_reward):
                f.write(f"actions: {act},\n observations: {obs},\n total reward: {sum_rew}\n")
        with open(os.path.join(log_path, "actions.txt"), "w") as f:
            f.write(version_header + "\n" + author_header + "\n" + reward_string + "\n")
            for act in zip(actions_log):
                f.write(f"actions: {act}")
        with open(os.path.join(log_path, "summary.json"), "w") as f:
            json.dump({"submission": {"author": getattr(submission, "NAME", None), "team": getattr(submission, "TEAM", None),
                                      "technique": getattr(submission, "TECHNIQUE", None)},
                       "parameters": {"seed": seed, "episode_length": EPISODE_LENGTH, "max_episodes": max_eps},
                       "time": {"start": str(start), "end": str(end), "elapsed": str(end - start)},
                       "reward": {"mean": out["mean"], "stdev": out["stdev"]},
                       "agents": {agent: str(submission.AGENTS[agent]) for agent in submission.AGENTS}}, f)
        with open(os.path.join(log_path, "scores.txt"), "w") as f:
            f.write(f"reward_mean: {out['mean']}\nreward_stdev: {out['stdev']}\n")
    return out


class MLPPolicy(torch.nn.Module):
    """BASELINE config 5: one random-init MLP per blue agent (obs -> 256 -> 256 -> n_actions), masked argmax."""

    cuda_graph_capturable = True        # static shapes, no host round trip, no private generator: run_evaluation_batched may graph it

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
    """The reference's cc4BlueRandomAgent (CybORG/Agents/SimpleAgents/RandomAgent.py:28-56, 66-69) as a device-side batched policy
    over the wrapper's index space: every blue agent first draws the action CLASS uniformly among Sleep, Monitor, Analyse,
    Remove, DeployDecoy and Restore (BlockTrafficZone / AllowTrafficZone are its `excluded_actions`), then the class's
    parameter uniformly among the known entries: here the host, uniform over the agent's existing hosts (mask True).  So
    Sleep and Monitor come up one time in six each, as under the reference agent, not once per index.  One difference
    remains: the reference agent's hostname dict also holds the subnet routers, which BlueFixedActionWrapper indices cannot
    address (BlueFixedActionWrapper.py:290).  `policy(obs, mask) -> int32 [N, 5]`."""

    CLASSES = ("Sleep", "Monitor", "Analyse", "Remove", "DeployDecoy", "Restore")

    def __init__(self, seed: int = 0):
        self.seed, self._g = seed, None

    @torch.no_grad()
    def __call__(self, obs: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
        if self._g is None:
            self._g = torch.Generator(device=mask.device).manual_seed(self.seed)
        n = mask.shape[0]
        cls = torch.randint(0, 6, (n, 5), generator=self._g, device=mask.device)
        out = torch.empty((n, 5), dtype=torch.int64, device=mask.device)
        for b in range(5):
            H, T = (48, 24) if b == 4 else (16, 8)
            # index layout of BlueFixedActionWrapper._populate_action_space :233-309
            fixed = torch.tensor([3 * H + 1, H, 0, H + 1, 3 * H + 2 + 2 * T, 2 * H + 1], device=mask.device)     # Sleep, Monitor, base of Analyse / Remove / DeployDecoy / Restore
            valid = mask[:, b, :H]                                    # the host blocks all carry the same mask: the host exists
            host = torch.rand((n, H), generator=self._g, device=mask.device).masked_fill(~valid, -1.0).argmax(-1)
            c = cls[:, b]
            out[:, b] = fixed[c] + torch.where(c >= 2, host, torch.zeros_like(host))
        return out.to(torch.int32).contiguous()


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
                           steps: int = EPISODE_LENGTH, graph: bool | None = None, return_actions: bool = False) -> dict:
    """The protocol with every episode of this rank's shard stepped concurrently; `policy(obs, mask) -> int32 [N,5]`
    runs on device and nothing crosses to the host during the episode: every episode ends on tick steps - 1
    (EnterpriseScenarioGenerator.determine_done :862-871), whose reward the protocol drops (evaluation.py:104-110), so the loop
    is steps - 2 counted ticks plus the final one, with no per-tick `done` read.  graph=True records one tick (policy, step,
    accumulation) into a CUDA graph and replays it: a 100-episode batch is launch-bound (about twenty small kernels per tick).
    Only a policy that is safe to capture may be graphed (static shapes, no host synchronisation, no private torch.Generator):
    the default (graph=None) graphs a policy that says so with a true `cuda_graph_capturable` attribute (MLPPolicy does) and runs
    every other policy eagerly; a capture that raises falls back to eager ticks.
    return_actions=True adds the submitted actions int32 [ticks, N, 5] (tests replay them through the oracle)."""
    lo, hi = shard(max_eps, rank, world)
    n = hi - lo
    dev = torch.device("cuda", device)
    totals = torch.zeros(max(n, 0), dtype=torch.float64, device=dev)
    t0 = time.time()
    log = None
    if n > 0:
        env = BatchedCC4(n, device=device, seed=seed, steps=steps, first_episode=lo)
        obs, mask = env.reset()
        acts = torch.zeros((n, 5), dtype=torch.int32, device=dev)
        counted = steps - 2                   # ticks 0 .. steps - 3 count; tick steps - 2 sets done and its reward is dropped
        log = torch.zeros((counted + 1, n, 5), dtype=torch.int32, device=dev) if return_actions else None

        def tick(count: bool):
            acts.copy_(policy(env.obs, mask))
            env.step(acts)
            if count:
                totals.add_(env.reward)

        g = None
        if graph is None:
            graph = bool(getattr(policy, "cuda_graph_capturable", False))
        if graph and counted > 2 and log is None:
            tick(True)                        # warm-up outside the capture (lazy allocations, cuBLAS handles); it is tick 0
            done_ticks = 1
            try:
                torch.cuda.synchronize(dev)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    tick(True)
            except Exception:               # noqa: BLE001  (policy not capturable)
                g = None
                torch.cuda.synchronize(dev)
                env = BatchedCC4(n, device=device, seed=seed, steps=steps, first_episode=lo)       # start over, eagerly
                obs, mask = env.reset()
                totals.zero_()
                done_ticks = 0
            if g is not None:
                for _ in range(counted - done_ticks):
                    g.replay()
                done_ticks = counted
        else:
            done_ticks = 0
        for t in range(done_ticks, counted):
            tick(True)
            if log is not None:
                log[t] = acts
        tick(False)
        if log is not None:
            log[counted] = acts
        assert bool(env.done.all()), "episodes must end on tick steps - 1"
    import torch.distributed as dist
    if world > 1 and not (dist.is_available() and dist.is_initialized()):
        # a shard evaluated on its own (no process group): its totals only; the caller combines them (aggregate_scores needs the group)
        vals = totals.cpu().tolist()
        out = {"mean": mean(vals) if vals else 0.0, "stdev": stdev(vals) if len(vals) > 1 else 0.0, "episodes": len(vals), "rewards": vals}
    else:
        out = aggregate_scores(totals, max_eps, rank, world)
    out["local_rewards"] = totals.cpu().tolist()
    out["seconds"] = time.time() - t0
    if return_actions:
        out["actions"] = log
    return out
