#!/usr/bin/env python
"""Where does a tick's time go?  Per-phase busy / barrier-wait cycles of every warp (= episode), from a profiling
build of the engine (-DCC4_PHASE_CLOCKS adds clock64() reads around the block barriers of the tick).

    python tools/phase_clocks.py --build          # here: nvcc the profiling variant into cage_challenge_4_b200/variants/
    python tools/phase_clocks.py [--envs 65536 --warm 150 --ticks 50]     # on the GPU box

Phases (cc4_engine.cu, enum PH_*): 0 load + tick setup, 1 policies + duration queue, 2 filter + blue actions, 3 green
agents + phishing, 4 red actions, 5 session scan + reassignment, 6 Monitor + RedSessionCheck + FSM digest, 7 counters /
reward + observation vectors + store.  "busy" = cycles the warp that took an episode spent on it; "wall" = cycles of the
whole block from the start of the phase to past its barrier; warp util = busy x episodes / (warps x wall).
"""
import ctypes as C, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "cage_challenge_4_b200", "variants", "libcc4_b200_phaseclk.so")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared", "-Xcompiler", "-fPIC", "-DCC4_PHASE_CLOCKS"]

if "--build" in sys.argv:
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    subprocess.check_call(["nvcc", *FLAGS, *[a for a in sys.argv[1:] if a.startswith("-D")], "-o", LIB, os.path.join(ROOT, "cage_challenge_4_b200", "csrc", "cc4_engine.cu")])
    print(LIB)
    sys.exit(0)

os.environ["CC4_B200_LIB"] = LIB
sys.path.insert(0, ROOT)
import numpy as np, torch
from cage_challenge_4_b200 import BatchedCC4, _native

def arg(name, default):
    return int(sys.argv[sys.argv.index(name) + 1]) if name in sys.argv else default

n, warm, ticks, wpb = arg("--envs", 65536), arg("--warm", 150), arg("--ticks", 50), arg("--wpb", 32)
env = BatchedCC4(n, device=0, seed=0)
env.reset()
g = torch.Generator(device="cuda").manual_seed(0)
def acts():
    return torch.stack([torch.randint(0, k, (n,), generator=g, device="cuda") for k in (82, 82, 82, 82, 242)], 1).to(torch.int32).contiguous()
for _ in range(warm):
    env.step(acts())
buf = np.zeros((n, 18), dtype=np.uint64)
_native.check(env._L.cc4_phase_clocks(env._h, buf.ctypes.data, 1), "cc4_phase_clocks")     # arms the counters
for _ in range(ticks):
    env.step(acts())
_native.check(env._L.cc4_phase_clocks(env._h, buf.ctypes.data, 1), "cc4_phase_clocks")
busy = buf[:, 0:16:2].astype(np.float64) / ticks        # per episode: cycles the warp that took it spent on the phase
wall = buf[:, 1:16:2].astype(np.float64) / ticks        # per block (slot of its first episode): phase start -> past the barrier
blocks = wall.sum(1) > 0
nblk = int(blocks.sum())
epb = n / max(1, nblk)
names = ["load+setup", "policies", "blue", "green", "red", "reassign", "end turn", "finish+obs+store"]
print(f"{n} episodes in {nblk} blocks ({epb:.1f} per block, {wpb} warps), {ticks} ticks after {warm} warm-up ticks")
print(f"block wall per tick {wall[blocks].sum(1).mean():.0f} cycles; episode busy per tick {busy.sum(1).mean():.0f} cycles")
print(f"{'phase':>18} {'busy/ep':>9} {'p90':>8} {'p99':>8} {'max':>8} {'wall/blk':>9} {'wall%':>6} {'warp util':>9}")
for i in range(8):
    b = busy[:, i]
    w = wall[blocks, i].mean()
    print(f"{names[i]:>18} {b.mean():9.0f} {np.percentile(b, 90):8.0f} {np.percentile(b, 99):8.0f} {b.max():8.0f} {w:9.0f} "
          f"{100 * w / wall[blocks].sum(1).mean():6.1f} {b.mean() * epb / (wpb * w):9.2f}")
ac = np.zeros((64, 2), dtype=np.uint64)
_native.check(env._L.cc4_action_clocks(env._h, ac.ctypes.data), "cc4_action_clocks")
names_a = {1: "Monitor", 3: "Remove", 4: "Restore", 5: "DeployDecoy", 10: "DiscoverRemoteSystems", 11: "AggressiveServiceDiscovery", 12: "StealthServiceDiscovery",
           13: "DiscoverDeception", 14: "ExploitRemoteService", 15: "PrivilegeEscalate", 16: "Impact", 17: "DegradeServices"}
print("executed actions (per episode-tick count, mean cycles each, share of the per-episode action time):")
tot_a = float(ac[:, 0].sum())
for code in sorted(names_a):
    if ac[code, 1]:
        print(f"  {names_a[code]:28s} {ac[code, 1] / (n * ticks):6.3f} x {ac[code, 0] / ac[code, 1]:7.0f} cycles  {100 * ac[code, 0] / tot_a:5.1f} %")
mx = buf[:, 17].astype(np.float64)
print(f"red phase: slowest single episode-tick per episode over the {ticks} ticks: mean {mx.mean():.0f}, p50 {np.percentile(mx, 50):.0f}, p99 {np.percentile(mx, 99):.0f}, max {mx.max():.0f} cycles; slowest in a block (over all its episodes and ticks): mean {buf[blocks, 16].astype(np.float64).mean():.0f}")
print(json.dumps({"busy_mean": busy.mean(0).tolist(), "block_wall_mean": wall[blocks].mean(0).tolist()}))
