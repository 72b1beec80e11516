"""e2e diagnostics (GPU box): device-only tick, host-buffer tick (bit rows / byte rows), a naive one-copy pipeline, and the host timeline
of a few calls (CC4_HOST_TRACE=1).  python tools/e2e_diag.py [n_envs]"""
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from cage_challenge_4_b200 import BatchedCC4

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
env = BatchedCC4(n, device=0, seed=0); env.reset()
g = torch.Generator().manual_seed(0)
acts = [torch.stack([torch.randint(0, k, (n,), generator=g) for k in (82, 82, 82, 82, 242)], 1).to(torch.int32).pin_memory() for _ in range(4)]
dacts = [a.cuda() for a in acts]

def timed(label, f, warm=60, ticks=300):
    env.reset()          # every section covers the same ticks of the episode (a tick gets more expensive as the episode goes on)
    for i in range(warm): f(i)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(ticks): f(i)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / ticks
    print(f"{label:40s} {dt * 1e3:.4f} ms per tick  {5 * n / dt / 1e6:.0f} M agent-steps/s", flush=True)

timed("device step_compact (async launches)", lambda i: env.step_compact(dacts[i % 4]))
def dev_sync(i):
    env.step_compact(dacts[i % 4]); torch.cuda.synchronize()
timed("device step_compact + sync per tick", dev_sync)
def dev_bits_sync(i):
    env.step_bits(dacts[i % 4]); torch.cuda.synchronize()
timed("device step_bits + sync per tick", dev_bits_sync)
timed("step_host bits", lambda i: env.step_host(acts[i % 4], bits=True))
timed("step_host compact (byte rows)", lambda i: env.step_host(acts[i % 4], compact=True))
if os.environ.get("CC4_HOST_TRACE"):
    os.environ["CC4_HOST_TRACE_ON"] = "1"
    env.reset()
    for i in range(200): env.step_host(acts[i % 4], compact=True)
