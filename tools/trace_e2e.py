import sys, os, time
sys.path.insert(0, os.getcwd())
import torch, numpy as np
from cage_challenge_4_b200 import BatchedCC4
n = 65536
env = BatchedCC4(n, device=0, seed=0); env.reset()
g = torch.Generator().manual_seed(0)
acts = [torch.stack([torch.randint(0, k, (n,), generator=g) for k in (82, 82, 82, 82, 242)], 1).to(torch.int32).pin_memory() for _ in range(4)]
for i in range(150): env.step_host(acts[i % 4], compact=True)
t0 = time.perf_counter()
for i in range(20): env.step_host(acts[i % 4], compact=True)
print("ms per tick", (time.perf_counter() - t0) / 20 * 1e3, file=sys.stderr)
