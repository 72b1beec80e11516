#!/usr/bin/env python
"""Where does a tick's time go?  Per-phase busy / barrier-wait cycles of every warp (= episode), from a profiling
build of the engine (-DCC4_PHASE_CLOCKS adds clock64() reads around the block barriers of the tick).

    python tools/phase_clocks.py --build          # here: nvcc the profiling variant into cage_challenge_4_b200/variants/
    python tools/phase_clocks.py [--envs 65536 --warm 150 --ticks 50]     # on the GPU box

Phase i = the code between block barrier i-1 and barrier i of w_tick (cc4_engine.cu): 0 tick setup, 1 policies +
duration queue, 2 filter + blue actions, 3 green agents + phishing, 4 red actions, 5 session scan + reassignment,
6 Monitor + RedSessionCheck + FSM digest, 7 counters / reward (the observation writer and the store follow barrier 7).
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

n, warm, ticks, wpb = arg("--envs", 65536), arg("--warm", 150), arg("--ticks", 50), arg("--wpb", 16)
env = BatchedCC4(n, device=0, seed=0)
env.reset()
g = torch.Generator(device="cuda").manual_seed(0)
def acts():
    return torch.stack([torch.randint(0, k, (n,), generator=g, device="cuda") for k in (82, 82, 82, 82, 242)], 1).to(torch.int32).contiguous()
for _ in range(warm):
    env.step(acts())
buf = np.zeros((n, 8, 2), dtype=np.uint64)
_native.check(env._L.cc4_phase_clocks(env._h, buf.ctypes.data, 1), "cc4_phase_clocks")     # arms the counters
for _ in range(ticks):
    env.step(acts())
_native.check(env._L.cc4_phase_clocks(env._h, buf.ctypes.data, 1), "cc4_phase_clocks")
busy, wait = buf[:, :, 0].astype(np.float64) / ticks, buf[:, :, 1].astype(np.float64) / ticks
nb = n // wpb
bb = busy[:nb * wpb].reshape(nb, wpb, 8)
tot = (busy + wait).sum(1).mean()
print(f"{n} episodes, {ticks} ticks after {warm} warm-up ticks; cycles per tick per warp (mean {tot:.0f} inside the barriers)")
print(f"{'phase':>5} {'busy':>8} {'wait':>8} {'blkmax':>8} {'p50':>8} {'p90':>8} {'p99':>8} {'max':>8}  busy% wait%")
for i in range(8):
    b = busy[:, i]
    print(f"{i:>5} {b.mean():8.0f} {wait[:, i].mean():8.0f} {bb[:, :, i].max(1).mean():8.0f} {np.percentile(b, 50):8.0f} {np.percentile(b, 90):8.0f} "
          f"{np.percentile(b, 99):8.0f} {b.max():8.0f}  {100 * b.mean() / tot:5.1f} {100 * wait[:, i].mean() / tot:5.1f}")
print(json.dumps({"busy_mean": busy.mean(0).tolist(), "wait_mean": wait.mean(0).tolist(), "block_max_busy_mean": bb.max(1).mean(0).tolist()}))
