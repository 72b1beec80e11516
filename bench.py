#!/usr/bin/env python
"""bench.py — agent-steps/s of the batched CC4 step (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--envs E]

A "step" is one tick (SimulationController.step) of every episode of the batch: `--envs` episodes per GPU (default
65 536, BASELINE config "65536 batched envs on 1xB200, full step incl. per-agent obs vector + BlueRewardMachine"),
FiniteStateRed + EnterpriseGreen scripted in-kernel, uniform-random blue indices and zero messages as inputs.
`value` = 5 blue agents x episodes x K / device time (CUDA events, max over ranks), inputs resident in HBM;
`e2e` = the same through the public host-buffer call (BatchedCC4.step_host -> cc4_step_host): H2D of the actions,
kernel, D2H of observations + reward + done inside the timed region.
`--impl reference` times the CPU restatement of the reference (oracle/, kind "port": the Python reference itself cannot
travel to the GPU box) on all host cores, same workload and metric.
One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_BLUE = 5
B_IO = 20 + 5 + (4 * 92 + 210) + 4 + 1      # actions + messages + 5 observation vectors + reward + done
SVC_BYTES = 128 * 8                          # per-host service table: read in place every tick by the green agents (not staged)
EPISODE_TICKS = 499                          # steps=500 -> done at step_count >= 499 (EnterpriseScenarioGenerator.py:870)
# dram__bytes_read.sum + dram__bytes_write.sum of one k_step launch at 65536 episodes from the committed
# `ncu --set full` capture (profiles/), or None when no capture matches the current kernel
NCU_TRAFFIC_BYTES_PER_LAUNCH = 462_498_560      # profiles/r02v_kstep_ncu_full_summary.txt: 301.0 MB read + 161.5 MB written


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.rows, self._halt = index, [], threading.Event()

    def run(self):
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=6)
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows for i in range(4) if len(r) > 2 + i and r[2 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


def cpu_port_throughput(seconds_target: float, threads: int):
    """The oracle port on `threads` host threads (ctypes releases the GIL): whole episodes with the same synthetic blue
    policy the GPU rollout uses.  Returns (env ticks per second, ticks, elapsed, episodes)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py
    oracle_py.lib()
    t0 = time.perf_counter()
    ticks, _ = oracle_py.run_episodes(0, 10_000_000, 1, 500)       # calibrate: one episode on one thread
    per_ep = time.perf_counter() - t0
    eps_per_thread = max(1, int(seconds_target / max(per_ep, 1e-4)))
    res = [0] * threads

    def work(i):
        res[i] = oracle_py.run_episodes(0, 1000 * i, eps_per_thread, 500)[0]

    ths = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    t0 = time.perf_counter()
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    dt = time.perf_counter() - t0
    return sum(res) / dt, sum(res), dt, eps_per_thread * threads


def run_reference(args, rank):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    vals = []
    for _ in range(max(1, args.warmup and 1)):
        cpu_port_throughput(1.0, threads)
    t_budget = min(20.0, max(2.0, 60.0 / max(1, args.steps)))
    ticks = dt = eps = 0
    for _ in range(max(1, min(args.steps, 5))):
        v, tk, d, ep = cpu_port_throughput(t_budget, threads)
        vals.append(v)
        ticks, dt, eps = ticks + tk, dt + d, eps + ep
    env_sps = ticks / dt
    value = env_sps * N_BLUE
    sample = f"{eps} whole 499-tick episodes (reset included) of the same workload, {threads} threads, {dt:.1f}s"
    line = {
        "impl": "reference", "metric": "agent-steps/sec", "value": value, "unit": "agent-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * args.envs * args.gpus / env_sps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": "agent-steps/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "env_steps_per_s": env_sps,
        "note": "CPU restatement (oracle/cc4_oracle.c) of the Python reference; the unmodified Python reference measured "
                "20.8 env-steps/s on one core of the build container (BASELINE.md)",
    }
    print(json.dumps(line), flush=True)


def workload_config(args):
    return {"workload": f"{args.envs} CC4 episodes per GPU (EnterpriseScenarioGenerator steps=500), FiniteStateRed + "
                        f"EnterpriseGreen in-kernel, uniform-random blue indices, full tick incl. 5 observation vectors + "
                        f"BlueRewardMachine (BASELINE configs[2])",
            "envs_per_gpu": args.envs, "episode_steps": 500, "blue_policy": "uniform random index (masked-out = Sleep)",
            "observations": "uint8 [envs, 578] per tick = the five blue vectors as BlueFlatWrapper returns them without pad_spaces (92 x 4 + 210), written to HBM",
            "cache": "per-tick state traffic (envs x record) is far larger than the 126 MB L2, no flush needed",
            "parallelism": f"episodes sharded over {args.gpus} GPU(s), no data-path collective"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=499)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs", type=int, default=65536, help="episodes per GPU")
    ap.add_argument("--e2e-steps", type=int, default=499)     # one whole episode through the host-buffer API
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    from cage_challenge_4_b200 import BatchedCC4
    from cage_challenge_4_b200.engine import STEP_ALL

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the engine has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # NCCL announces its version on STDOUT when the communicator is created; stdout carries the one JSON line only,
        # so the communicator is created (init + first collective) with fd 1 pointing at stderr
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    n = args.envs
    env = BatchedCC4(n, device=local_rank, seed=0, steps=500, first_episode=rank * n)
    dyn_bytes, cold_bytes, stat_bytes = env.state_bytes()
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    # one fresh uniform-random action tensor per tick (a short repeating pool would make every episode re-issue the same
    # few actions, e.g. dozens of decoys on one host, which is not the workload BASELINE names)
    n_pool = min(max(args.steps, args.e2e_steps, 16), 512)
    pool = [torch.stack([torch.randint(0, k, (n,), generator=g, device=dev, dtype=torch.int32) for k in (82, 82, 82, 82, 242)], 1).contiguous()
            for _ in range(n_pool)]
    ret = torch.zeros(n, dtype=torch.float64, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    tick = 0

    def one_step(i):
        nonlocal tick
        if tick == EPISODE_TICKS:            # every episode of the batch is done: reset (inside the timed region)
            env.reset()
            tick = 0
        obs, rew, done = env.step_compact(pool[i % len(pool)], None)     # obs uint8 [N, 578]: the five unpadded vectors (BlueFlatWrapper default, pad_spaces=False)
        ret.add_(rew)
        tick += 1

    env.reset()
    for i in range(max(3, args.warmup)):
        one_step(i)
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    r0.record()
    env.reset()                              # k_reset of the whole batch, timed on its own (reported as reset_ms)
    r1.record()
    torch.cuda.synchronize()
    reset_ms = r0.elapsed_time(r1)
    tick = 0
    ret.zero_()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = env.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # the dominant kernel (k_step) is also timed alone, launch by launch, with CUDA events on the launching stream
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    e0.record()
    for i in range(args.steps):
        if tick == EPISODE_TICKS:            # every episode of the batch is done: reset (inside the timed region)
            env.reset()
            tick = 0
        kev[i][0].record()
        obs, rew, done = env.step_compact(pool[i % len(pool)], None)     # obs uint8 [N, 578]: the five unpadded vectors (BlueFlatWrapper default, pad_spaces=False)
        kev[i][1].record()
        ret.add_(rew)
        tick += 1
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = env.launch_count() - l0
    kernel_ms = sum(a.elapsed_time(b) for a, b in kev) / args.steps

    # end to end through the public host-buffer API: pinned actions in, pinned obs / reward / done out
    host_pool = [p.cpu().pin_memory() for p in pool[:max(args.e2e_steps, 4)]]
    env.reset()
    for i in range(3):
        env.step_host(host_pool[i % len(host_pool)], compact=True)
    env.reset()
    torch.cuda.synchronize()
    barrier()
    tick = 0
    t0 = time.perf_counter()
    for i in range(args.e2e_steps):
        if tick == EPISODE_TICKS:            # the batch has finished its episode: reset inside the timed region
            env.reset()
            tick = 0
        o, r, d = env.step_host(host_pool[i % len(host_pool)], compact=True)    # obs uint8 [N, 578]: the five unpadded vectors
        tick += 1
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop()      # sampled through both timed regions (device-resident steps, then the host-buffer steps)
    h2d = n * N_BLUE * 4
    expand = os.environ.get("CC4_HOST_EXPAND", "1")[:1] != "0"
    d2h = n * ((80 if expand else 578) + 4 + 1)     # observation rows cross PCIe as 80-byte bit rows and are expanded on the host

    t_ms = torch.tensor([ms, e2e_s * 1e3, kernel_ms], dtype=torch.float64, device=dev)
    score = torch.stack([ret.sum(), (ret * ret).sum(), torch.tensor(float(n), dtype=torch.float64, device=dev)])
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(score, op=dist.ReduceOp.SUM)    # the only collective on the path: final score aggregation
    ms, e2e_ms, kernel_ms = (float(x) for x in t_ms.tolist())
    errs = env.error_flags()

    if rank == 0:
        peak, which = load_peaks()
        total_envs = n * world
        env_sps = total_envs * args.steps / (ms / 1e3)
        b_step = 2 * dyn_bytes + SVC_BYTES + B_IO
        achieved = b_step * n / (kernel_ms / 1e3) / 1e9
        line = {
            "metric": "agent-steps/sec", "value": env_sps * N_BLUE, "unit": "agent-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic", "config": workload_config(args),
            "env_steps_per_s": env_sps, "all_agent_steps_per_s": env_sps * 70.0,
            # an episode is 499 ticks and one reset of the batch (EnterpriseScenarioGenerator on device); the timed region holds a
            # reset whenever it crosses an episode boundary (--steps > 499); the same figure with one reset per 499 ticks charged:
            "reset_ms": reset_ms,
            "value_incl_reset": total_envs * N_BLUE * EPISODE_TICKS / ((ms / args.steps * EPISODE_TICKS + reset_ms) / 1e3),
            "gpu_launches": int(launches), "clocks": clocks,
            "e2e": {"value": total_envs * args.e2e_steps * N_BLUE / (e2e_ms / 1e3), "unit": "agent-steps/s",
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": args.e2e_steps,
                    "api": "BatchedCC4.step_host(actions_pinned, compact=True) -> cc4_step_host: obs uint8 [N,578] (92 x 4 + 210, the reference's unpadded vectors) + reward f32 [N] + done u8 [N] in pinned host memory",
                    "obs_over_pcie": "bit rows uint32 [N,20], expanded to the uint8 rows by cc4_step_host's host threads inside the timed region" if expand else "uint8 rows"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": NCU_TRAFFIC_BYTES_PER_LAUNCH if n == 65536 else None, "peak_source": which, "kernel": "k_step",
                         "kernel_ms": kernel_ms, "bytes_per_env_step": b_step,
                         "formula": "2*sizeof(hot record) + 1024 B service table (read in place) + 608 B io per episode-tick; the rest of the cold record is touched by rare transitions only (DESIGN.md); traffic = ncu dram bytes of one launch at 65536 episodes"},
            "state_bytes": {"hot": dyn_bytes, "cold": cold_bytes, "static": stat_bytes},
            "score": {"episodes": int(score[2].item()), "mean_return": float(score[0] / score[2])},
            "error_episodes": int((errs != 0).sum()),
        }
        if world == 1 and not args.no_cpu:
            threads = 1
            v, tk, dt, eps = cpu_port_throughput(args.cpu_seconds, threads)
            line["cpu_baseline"] = {"value": v * N_BLUE, "unit": "agent-steps/s", "cores": threads, "kind": "port",
                                    "sample": f"{eps} whole 499-tick episodes of the same workload on 1 host thread, {dt:.1f}s "
                                              f"(oracle/cc4_oracle.c; the Python reference itself: 20.8 env-steps/s/core, BASELINE.md)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
