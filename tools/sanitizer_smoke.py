#!/usr/bin/env python
"""Small run of every output path for compute-sanitizer (run on the GPU box):
    compute-sanitizer --tool memcheck python tools/sanitizer_smoke.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cage_challenge_4_b200 import BatchedCC4
from cage_challenge_4_b200.engine import unpack_bits

n = 200
envs = [BatchedCC4(n, device=0, seed=9) for _ in range(4)]
for e in envs:
    e.reset()
g = np.random.default_rng(0)
for t in range(60):
    acts = torch.from_numpy(np.stack([g.integers(0, k, n) for k in (82, 82, 82, 82, 242)], 1).astype(np.int32))
    msgs = torch.from_numpy(g.integers(0, 256, (n, 5)).astype(np.uint8))
    o0, r0, d0 = envs[0].step(acts.cuda(), msgs.cuda())
    o1, r1, d1 = envs[1].step_compact(acts.cuda(), msgs.cuda())
    o2, r2, d2 = envs[2].step_bits(acts.cuda(), msgs.cuda())
    o3, r3, d3 = envs[3].step_host(acts.pin_memory(), msgs.pin_memory(), compact=True)
    assert torch.equal(unpack_bits(o2), o1) and torch.equal(o3, o1.cpu()) and torch.equal(r0, r2)
envs[0].rollout(40, policy_seed=3)
# the other scripted red agents (their own instantiations of k_step), bit rows to the host, a second reset (k_reset_fast + k_reset_outputs)
for red in ("discovery", "random_select"):
    e = BatchedCC4(150, device=0, seed=4, red_agent=red)
    e.reset()
    for t in range(80):
        acts = torch.from_numpy(np.stack([g.integers(0, k, 150) for k in (82, 82, 82, 82, 242)], 1).astype(np.int32))
        if t % 3:
            acts[:] = torch.tensor([49, 49, 49, 49, 145], dtype=torch.int32)
        e.step_host(acts.pin_memory(), bits=True)
    e.reset()
    e.step_state_only(acts.cuda())
    envs.append(e)
torch.cuda.synchronize()
assert not any(e.error_flags().any() for e in envs)
print("sanitizer smoke ok")
