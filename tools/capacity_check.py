import sys, numpy as np, torch
import os; ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from cage_challenge_4_b200 import BatchedCC4
n = 65536
env = BatchedCC4(n, device=0, seed=123)
env.reset()
acts = torch.tensor([[49, 49, 49, 49, 145]], dtype=torch.int32, device="cuda").repeat(n, 1).contiguous()
for t in range(499):
    env.step(acts)
torch.cuda.synchronize()
e = env.error_flags()
print("all-Sleep blue, 65536 episodes x 499 ticks: episodes with error flags:", int((e != 0).sum()), "capacity:", int(((e & 4) != 0).sum()), "unreachable:", int(((e & 8) != 0).sum()))
# session counts from the canonical export of a sample
mx = 0
for i in range(0, 400):
    c = env.export_canon(i)
    import importlib; sys.path.insert(0, os.path.join(ROOT, 'oracle')); import canon
    st = canon.decode(list(c))
    mx = max(mx, sum(len(r["sessions"]) for r in st["red"]))
print("max total red sessions in 400 sampled episodes:", mx)
