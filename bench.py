#!/usr/bin/env python
"""bench.py — agent-steps/s of the batched CC4 step (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--envs E]

A "step" is one tick (SimulationController.step) of every episode of the batch: `--envs` episodes per GPU (default
65 536, BASELINE config "65536 batched envs on 1xB200, full step incl. per-agent obs vector + BlueRewardMachine"),
FiniteStateRed + EnterpriseGreen scripted in-kernel, uniform-random blue indices and zero messages as inputs.
`value` = 5 blue agents x episodes x K / device time (CUDA events, max over ranks), inputs resident in HBM;
`e2e` = the same through the public host-buffer call (BatchedCC4.step_host -> cc4_step_host): H2D of the actions,
kernel, D2H of observations + reward + done inside the timed region.
`configs` carries the other BASELINE configs: cfg2 (4 096 episodes, state transition only: per-tick launches and the fused
rollout), cfg4 (65 536 episodes in TOTAL, split over the ranks: strong scaling; per-tick launches and a CUDA graph of them),
cfg5 (the evaluation protocol: 100 x 500-step episodes under a random-init MLP blue policy on device).
`--impl reference` times the UNMODIFIED Python reference (baseline/_ref, one process per host core, same workload through
BlueFlatWrapper.step) and, beside it, the C restatement (oracle/, kind "port"); without baseline/_ref it falls back to the
port and says so.
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
N_SM, ISSUE_PER_SM = 148, 4                  # B200: 148 SMs, 4 warp schedulers each, one warp-instruction per scheduler and cycle
# From the committed `ncu --set full` capture of one k_step launch at 65 536 episodes (profiles/, see NCU_SOURCE):
# dram__bytes_read.sum + dram__bytes_write.sum, and smsp__inst_executed.sum (warp-instructions).
NCU_SOURCE = "profiles/r05w_kstep_ncu_full_summary.txt"
NCU_TRAFFIC_BYTES_PER_LAUNCH = 456_952_064      # 299.90 MB read + 157.05 MB written
NCU_WARP_INST_PER_LAUNCH = 203_514_083


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            j = json.load(f)
            return float(j["hbm_gbs"]), "measured", float(j.get("sm_max_mhz", 1965.0))
    return 6650.0, "fallback", 1965.0


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


# ---------------------------------------------------------------------------------------------------------------------
# CPU legs (the only places that execute oracle/ or baseline/_ref; never inside a GPU timed region)
# ---------------------------------------------------------------------------------------------------------------------
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


def reference_pool(procs=None):
    """Worker processes holding one unmodified-reference environment each (baseline/ref_runner.py), or None when
    baseline/_ref is absent.  Started BEFORE torch / CUDA are initialised in this process (the workers are forks)."""
    try:
        from baseline import ref_runner
        if not ref_runner.available():
            return None
        return ref_runner.ReferencePool(procs=procs, policy="random", seed_base=0)
    except Exception as e:  # noqa: BLE001
        print(f"bench.py: the Python reference is unavailable ({e}); CPU legs use the C port only", file=sys.stderr)
        return None


def run_reference(args, rank):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    pool = reference_pool(threads)
    samples = max(1, min(args.steps, 20))
    t_sample = min(20.0, max(1.5, 60.0 / samples))
    if pool is not None:
        for _ in range(min(args.warmup, 2)):
            pool.sample(1.0)
        ticks = wall = resets = 0
        rates = []
        for _ in range(samples):
            r, tk, w, rs = pool.sample(t_sample)
            rates.append(r)
            ticks, wall, resets = ticks + tk, wall + w, resets + rs
        pool.close()
        env_sps = sum(rates) / len(rates)
        kind = "reference"
        sample = (f"{ticks} ticks of the UNMODIFIED Python reference (baseline/_ref: CybORG 4.0 through BlueFlatWrapper.step, uniform-random blue "
                  f"indices, full tick incl. 5 observation vectors + BlueRewardMachine, {resets} episode resets inside), one process per host "
                  f"thread x {threads}, {samples} samples of {t_sample:.1f}s")
    else:
        kind, ticks, wall = "port", 0, 0.0
        for _ in range(samples if samples < 5 else 5):
            v, tk, d, ep = cpu_port_throughput(min(t_sample, 10.0), threads)
            ticks, wall = ticks + tk, wall + d
        env_sps = ticks / wall
        sample = f"C restatement (oracle/cc4_oracle.c) on {threads} threads, {ticks} ticks in whole 499-tick episodes; baseline/_ref is absent on this box"
    value = env_sps * N_BLUE
    # the C port on all threads beside it (short): how far the restatement is from the code it restates
    pv, ptk, pdt, pep = cpu_port_throughput(4.0, threads)
    line = {
        "impl": "reference", "metric": "agent-steps/sec", "value": value, "unit": "agent-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * args.envs * args.gpus / env_sps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": "agent-steps/s", "cores": threads, "kind": kind, "sample": sample},
        "cpu_baseline_port": {"value": pv * N_BLUE, "unit": "agent-steps/s", "cores": threads, "kind": "port",
                              "sample": f"{pep} whole 499-tick episodes of the same workload (oracle/cc4_oracle.c), {threads} threads, {pdt:.1f}s"},
        "e2e": {"value": value, "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "env_steps_per_s": env_sps,
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


# ---------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=499)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs", type=int, default=65536, help="episodes per GPU")
    ap.add_argument("--e2e-steps", type=int, default=499)     # one whole episode through the host-buffer API
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the cfg2 / cfg4 / cfg5 measurements")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    # the reference workers are forked before CUDA exists in this process and sleep on a pipe until the GPU part is over
    ref_pool = reference_pool() if (world == 1 and not args.no_cpu) else None

    import torch
    import torch.distributed as dist
    from cage_challenge_4_b200 import BatchedCC4, _native

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

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(*vals):
        t = torch.tensor(list(vals), dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t.tolist()]

    def action_pool(n, count, seed):
        g = torch.Generator(device=dev).manual_seed(seed)
        return [torch.stack([torch.randint(0, k, (n,), generator=g, device=dev, dtype=torch.int32) for k in (82, 82, 82, 82, 242)], 1).contiguous()
                for _ in range(count)]

    # =================================================================================================================
    # headline: configs[2] on every GPU (weak scaling)
    # =================================================================================================================
    n = args.envs
    env = BatchedCC4(n, device=local_rank, seed=0, steps=500, first_episode=rank * n)
    dyn_bytes, cold_bytes, stat_bytes = env.state_bytes()
    # one fresh uniform-random action tensor per tick (a short repeating pool would make every episode re-issue the same
    # few actions, e.g. dozens of decoys on one host, which is not the workload BASELINE names)
    n_pool = min(max(args.steps, args.e2e_steps, 16), 512)
    pool = action_pool(n, n_pool, 1234 + rank)
    ret = torch.zeros(n, dtype=torch.float64, device=dev)
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
    reset_times = []
    for _ in range(3):                       # k_reset of the whole batch, timed on its own (reported as reset_ms: median of three)
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r0.record()
        env.reset()
        r1.record()
        torch.cuda.synchronize()
        reset_times.append(r0.elapsed_time(r1))
    reset_ms = sorted(reset_times)[1]
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
        obs, rew, done = env.step_compact(pool[i % len(pool)], None)
        kev[i][1].record()
        ret.add_(rew)
        tick += 1
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = env.launch_count() - l0
    kernel_ms = sum(a.elapsed_time(b) for a, b in kev) / args.steps
    score = torch.stack([ret.sum(), (ret * ret).sum(), torch.tensor(float(n), dtype=torch.float64, device=dev)])

    # ---- end to end through the public host-buffer API: pinned actions in, pinned obs / reward / done out ----
    host_pool = [p.cpu().pin_memory() for p in pool[:max(min(args.e2e_steps, len(pool)), 4)]]

    def e2e_run(steps, **kw):
        env.reset()
        for i in range(3):
            env.step_host(host_pool[i % len(host_pool)], **kw)
        env.reset()
        torch.cuda.synchronize()
        barrier()
        t, t0 = 0, time.perf_counter()
        for i in range(steps):
            if t == EPISODE_TICKS:           # the batch has finished its episode: reset inside the timed region
                env.reset()
                t = 0
            env.step_host(host_pool[i % len(host_pool)], **kw)
            t += 1
        torch.cuda.synchronize()
        return time.perf_counter() - t0

    e2e_s = e2e_run(args.e2e_steps, compact=True)              # obs uint8 [N, 578]: the five unpadded vectors
    e2e_bits_s = e2e_run(min(args.e2e_steps, 200), bits=True)  # declared variant: obs as bit rows int32 [N, 20], nothing expanded on the host
    clocks = sampler.stop()      # sampled through the timed regions above (device-resident steps, then the host-buffer steps)
    h2d = n * N_BLUE * 4
    expand = os.environ.get("CC4_HOST_EXPAND", "1")[:1] != "0"
    d2h = n * ((80 if expand else 578) + 4 + 1)     # observation rows cross PCIe as 80-byte bit rows and are expanded on the host
    ms, e2e_ms, e2e_bits_ms, kernel_ms = reduce_max(ms, e2e_s * 1e3, e2e_bits_s * 1e3, kernel_ms)
    if world > 1:
        dist.all_reduce(score, op=dist.ReduceOp.SUM)    # the only collective on the path: final score aggregation
    errs = env.error_flags()
    # host side of the uint8-row path: what the pool's threads can write per second (all ranks measure at the same time)
    import ctypes as C
    L = _native.load()
    lw = int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1)
    host_threads = int(os.environ.get("CC4_HOST_THREADS", "0") or 0) or max(1, min(8, (os.cpu_count() or 1) // max(1, lw)))
    stream_rows = lw * n * 578 > (64 << 20)
    gbs = C.c_double(0.0)
    barrier()
    # same footprint as the rows of one tick (one rank's 38 MB stay in the host's caches, eight ranks' 303 MB do not)
    L.cc4_host_write_bandwidth(max(n * 578, 1 << 20), host_threads, 1 if stream_rows else 0, 20, C.byref(gbs))
    (host_write_gbs,) = [-x for x in reduce_max(-gbs.value)]      # the slowest rank's share while all ranks write at once
    del env
    torch.cuda.empty_cache()

    configs = {}
    if not args.no_configs:
        from cage_challenge_4_b200.evaluation import MLPPolicy, run_evaluation_batched

        def time_ticks(e, acts, ticks, step):
            """`ticks` per-tick launches on a fresh episode; device time in ms (max over ranks is taken by the caller)."""
            e.reset()
            for i in range(3):
                step(acts[i % len(acts)])
            e.reset()
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for i in range(ticks):
                step(acts[i % len(acts)])
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b)

        # ---- cfg2: 4 096 episodes on one GPU, FiniteStateRed + random blue, observations / reward off ----
        if world == 1:
            n2, t2 = 4096, 498
            e2 = BatchedCC4(n2, device=local_rank, seed=0, steps=500)
            acts2 = action_pool(n2, 64, 99)
            ms_tick = time_ticks(e2, acts2, t2, e2.step_state_only) / t2
            gr = e2.capture_steps(acts2[:32], mode="state")
            e2.reset()
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(15):
                gr.replay()
            b.record()
            torch.cuda.synchronize()
            ms_graph = a.elapsed_time(b) / (15 * 32)
            e2.reset()
            e2.rollout(8, policy_seed=1, flags=0)
            e2.reset()
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            e2.rollout(t2, policy_seed=1, flags=0)          # ONE launch: the hot records stay in shared memory for all the ticks
            b.record()
            torch.cuda.synchronize()
            ms_roll = a.elapsed_time(b) / t2
            hot2, io2 = 2 * dyn_bytes, 25
            configs["cfg2"] = {
                "workload": "BASELINE configs[1]: 4096 episodes on 1 GPU, FiniteStateRed + uniform-random blue, observations and reward off (state transition only)",
                "envs": n2, "ticks": t2, "blocks_per_launch": (n2 + 63) // 64 if n2 > 148 * 28 else None,
                "step": {"ms_per_tick": ms_tick, "value": n2 * N_BLUE / (ms_tick / 1e3), "bytes_per_env_step": hot2 + SVC_BYTES + io2,
                         "api": "cc4_step(flags=0), one launch per tick from Python"},
                "step_graph": {"ms_per_tick": ms_graph, "value": n2 * N_BLUE / (ms_graph / 1e3), "api": "the same 32 launches replayed as one CUDA graph"},
                "rollout": {"ms_per_tick": ms_roll, "value": n2 * N_BLUE / (ms_roll / 1e3), "K": t2, "bytes_per_env_step": hot2 / t2 + SVC_BYTES + 4,
                            "api": "cc4_rollout: K fused ticks in one launch, on-device random blue policy, records resident in shared memory (B_step / K for the hot record)"},
                "unit": "agent-steps/s", "error_episodes": int((e2.error_flags() != 0).sum()),
            }
            del e2

        # ---- cfg4: 65 536 episodes in TOTAL, split over the ranks (strong scaling; 8 192 per GPU at N = 8) ----
        def cfg4_run(n4, label):
            e4 = BatchedCC4(n4, device=local_rank, seed=0, steps=500, first_episode=rank * n4)
            acts4 = action_pool(n4, 64, 77 + rank)
            t4 = 200
            ms_tick = time_ticks(e4, acts4, t4, e4.step_compact) / t4
            acc = torch.zeros(n4, dtype=torch.float64, device=dev)
            gr = e4.capture_steps(acts4[:20], mode="compact", accumulate=acc)
            e4.reset()
            acc.zero_()
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(10):
                gr.replay()
            b.record()
            torch.cuda.synchronize()
            ms_graph = a.elapsed_time(b) / 200
            ms_tick, ms_graph = reduce_max(ms_tick, ms_graph)
            sc = torch.stack([acc.sum(), torch.tensor(float(n4), dtype=torch.float64, device=dev)])
            if world > 1:
                dist.all_reduce(sc, op=dist.ReduceOp.SUM)       # NCCL score all-reduce (the only collective of the config)
            epb = -(-n4 // (-(-n4 // (N_SM * 64)) * N_SM))
            out = {"workload": label, "envs_total": n4 * world, "envs_per_gpu": n4, "n_gpus": world, "scaling": "strong", "ticks": 200,
                   "blocks_per_launch": -(-n4 // epb), "episodes_per_block": epb, "sub_wave": -(-n4 // epb) < N_SM,
                   "step": {"ms_per_tick": ms_tick, "value": n4 * world * N_BLUE / (ms_tick / 1e3), "api": "cc4_step per tick from Python"},
                   "step_graph": {"ms_per_tick": ms_graph, "value": n4 * world * N_BLUE / (ms_graph / 1e3), "api": "20 ticks per CUDA graph replay (BatchedCC4.capture_steps)"},
                   "value": n4 * world * N_BLUE / (min(ms_tick, ms_graph) / 1e3), "unit": "agent-steps/s",
                   "mean_return_200_ticks": float(sc[0] / sc[1]), "error_episodes": int((e4.error_flags() != 0).sum())}
            del e4
            return out

        if world > 1:
            configs["cfg4"] = cfg4_run(65536 // world, f"BASELINE configs[3]: 65536 episodes sharded over {world} GPUs ({65536 // world} per GPU), NCCL score all-reduce only")
        else:
            configs["cfg4_one_shard"] = cfg4_run(8192, "one GPU's share of BASELINE configs[3] (8192 of the 65536 episodes) measured on this GPU alone; the N-GPU line carries configs.cfg4")

        # ---- cfg5: the evaluation protocol, 100 x 500-step episodes, random-init MLP blue policy on device ----
        policy = MLPPolicy(seed=0).to(dev)
        run_evaluation_batched(policy, max_eps=100, seed=3, device=local_rank, rank=rank, world=world, steps=40)      # warm-up (allocations, cuBLAS handles)
        barrier()
        t0 = time.perf_counter()
        ev = run_evaluation_batched(policy, max_eps=100, seed=0, device=local_rank, rank=rank, world=world)
        torch.cuda.synchronize()
        (wall5,) = reduce_max(time.perf_counter() - t0)
        configs["cfg5"] = {"workload": f"BASELINE configs[4]: Evaluation/evaluation.py protocol, 100 x 500-step episodes vs FiniteStateRedAgent, random-init MLP blue policy "
                                       f"(obs -> 256 -> 256 -> actions, masked argmax) evaluated on device, episodes sharded over {world} GPU(s), one all_gather of the totals",
                           "episodes": ev["episodes"], "wall_s": wall5, "value": 100 * EPISODE_TICKS * N_BLUE / wall5, "unit": "agent-steps/s",
                           "mean": ev["mean"], "stdev": ev["stdev"], "api": "evaluation.run_evaluation_batched(MLPPolicy)"}

    if rank == 0:
        peak, which, sm_max = load_peaks()
        total_envs = n * world
        env_sps = total_envs * args.steps / (ms / 1e3)
        b_step = 2 * dyn_bytes + SVC_BYTES + B_IO
        achieved = b_step * n / (kernel_ms / 1e3) / 1e9
        sm_mhz = clocks.get("sm_mhz") or sm_max
        issue_peak = ISSUE_PER_SM * N_SM * sm_mhz * 1e6                 # warp-instructions per second the chip can issue at the sampled clock
        issue_ach = (NCU_WARP_INST_PER_LAUNCH * n / 65536) / (kernel_ms / 1e3)
        hbm_frac, issue_frac = achieved / peak, issue_ach / issue_peak
        e2e_value = total_envs * args.e2e_steps * N_BLUE / (e2e_ms / 1e3)
        e2e_bytes_per_s = e2e_value / N_BLUE * 578 / 1e9 / world        # uint8 row bytes this rank's pool writes per second
        line = {
            "metric": "agent-steps/sec", "value": env_sps * N_BLUE, "unit": "agent-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic", "config": workload_config(args),
            "env_steps_per_s": env_sps, "all_agent_steps_per_s": env_sps * 70.0,
            # an episode is 499 ticks and one reset of the batch (EnterpriseScenarioGenerator on device); the timed region holds a
            # reset whenever it crosses an episode boundary (--steps > 499); the same figure with one reset per 499 ticks charged:
            "reset_ms": reset_ms, "reset_ms_samples": reset_times,
            "value_incl_reset": total_envs * N_BLUE * EPISODE_TICKS / ((ms / args.steps * EPISODE_TICKS + reset_ms) / 1e3),
            "gpu_launches": int(launches), "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "agent-steps/s",
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": args.e2e_steps,
                    "api": "BatchedCC4.step_host(actions_pinned, compact=True) -> cc4_step_host: obs uint8 [N,578] (92 x 4 + 210, the reference's unpadded vectors) + reward f32 [N] + done u8 [N] in pinned host memory",
                    "obs_over_pcie": "bit rows uint32 [N,20], expanded to the uint8 rows by cc4_step_host's host threads inside the timed region" if expand else "uint8 rows",
                    "host_threads": host_threads, "host_write_gbs": host_write_gbs,
                    "host_write_note": f"{'streaming' if stream_rows else 'ordinary'} stores by {host_threads} threads per rank into a buffer of one tick's rows ({n * 578 >> 20} MB), all {world} rank(s) at once (cc4_host_write_bandwidth); "
                                       f"the uint8 rows this rank writes: {e2e_bytes_per_s:.1f} GB/s = {e2e_bytes_per_s / max(host_write_gbs, 1e-9):.2f} of that",
                    "host_write_frac": e2e_bytes_per_s / max(host_write_gbs, 1e-9)},
            "e2e_bits": {"value": total_envs * min(args.e2e_steps, 200) * N_BLUE / (e2e_bits_ms / 1e3), "unit": "agent-steps/s",
                         "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": n * (80 + 4 + 1), "steps": min(args.e2e_steps, 200),
                         "api": "BatchedCC4.step_host(actions_pinned, bits=True): obs as bit rows int32 [N,20] (CC4_STEP_OBS_BITS; engine.unpack_bits / cc4_expand_obs_bits "
                                "turn them into the uint8 rows) + reward + done in pinned host memory; declared variant of e2e with nothing to expand on the host"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": hbm_frac,
                         "traffic": int(NCU_TRAFFIC_BYTES_PER_LAUNCH * n / 65536) if n == 65536 else None, "peak_source": which, "kernel": "k_step",
                         "kernel_ms": kernel_ms, "bytes_per_env_step": b_step,
                         "formula": "2*sizeof(hot record) + 1024 B service table (read in place) + 608 B io per episode-tick; the rest of the cold record is touched by rare transitions only (DESIGN.md); traffic = ncu dram bytes of one launch at 65536 episodes",
                         # the path is branchy integer code: what actually binds it is warp-instruction ISSUE, reported beside the HBM figure
                         "binding": "issue" if issue_frac > hbm_frac else "hbm",
                         "issue": {"achieved": issue_ach / 1e9, "peak": issue_peak / 1e9, "unit": "G warp-inst/s", "frac": issue_frac,
                                   "warp_inst_per_launch": int(NCU_WARP_INST_PER_LAUNCH * n / 65536), "sm_mhz": sm_mhz,
                                   "formula": f"smsp__inst_executed.sum of one launch ({NCU_SOURCE}) / kernel_ms against {ISSUE_PER_SM} issue slots x {N_SM} SMs x the SM clock sampled during the run"}},
            "state_bytes": {"hot": dyn_bytes, "cold": cold_bytes, "static": stat_bytes},
            "score": {"episodes": int(score[2].item()), "mean_return": float(score[0] / score[2])},
            "error_episodes": int((errs != 0).sum()),
            "configs": configs,
        }
        if world == 1 and not args.no_cpu:
            if ref_pool is not None:
                ref_pool.sample(1.0)
                r, tk, w, rs = ref_pool.sample(args.cpu_seconds)
                line["cpu_baseline"] = {"value": r * N_BLUE, "unit": "agent-steps/s", "cores": ref_pool.procs, "kind": "reference",
                                        "sample": f"{tk} ticks of the UNMODIFIED Python reference (baseline/_ref, BlueFlatWrapper.step with uniform-random blue indices: the same "
                                                  f"full tick), one process per host thread x {ref_pool.procs}, {w:.1f}s, {rs} episode resets inside"}
                ref_pool.close()
            v, tk, dt, eps = cpu_port_throughput(min(args.cpu_seconds, 6.0), 1)
            port = {"value": v * N_BLUE, "unit": "agent-steps/s", "cores": 1, "kind": "port",
                    "sample": f"{eps} whole 499-tick episodes of the same workload on 1 host thread, {dt:.1f}s (oracle/cc4_oracle.c, the C restatement of the reference)"}
            if "cpu_baseline" in line:
                line["cpu_baseline_port"] = port
            else:
                line["cpu_baseline"] = port
        print(json.dumps(line), flush=True)
    elif ref_pool is not None:
        ref_pool.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
