#!/usr/bin/env python
"""Wall cycles of every phase of a tick, per block, from a -DCC4_PHASE_WALL build (one clock read by one thread after each
block barrier; does not perturb the kernel the way the per-item clocks of tools/phase_clocks.py do).

    python tools/phase_wall.py --build [name -Dflags...]     # here: nvcc the variant(s) into cage_challenge_4_b200/variants/
    python tools/phase_wall.py [name ...] [--envs 65536 --warm 150 --ticks 100]      # on the GPU box
"""
import os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VDIR = os.path.join(ROOT, "cage_challenge_4_b200", "variants")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared", "-Xcompiler", "-fPIC", "-DCC4_PHASE_WALL"]


def lib(name):
    return os.path.join(VDIR, f"libcc4_b200_wall_{name}.so")


if "--build" in sys.argv:
    os.makedirs(VDIR, exist_ok=True)
    rest = [a for a in sys.argv[1:] if a != "--build"]
    name = rest[0] if rest and not rest[0].startswith("-") else "base"
    subprocess.check_call(["nvcc", *FLAGS, *[a for a in rest if a.startswith("-D")], "-o", lib(name), os.path.join(ROOT, "cage_challenge_4_b200", "csrc", "cc4_engine.cu")])
    print(lib(name))
    sys.exit(0)


def arg(name, default):
    return int(sys.argv[sys.argv.index(name) + 1]) if name in sys.argv else default


names = [a for i, a in enumerate(sys.argv[1:], 1) if not a.startswith("--") and not sys.argv[i - 1].startswith("--")] or ["base"]
if len(names) > 1:      # one process per variant (the library is chosen at import time)
    for nm in names:
        subprocess.call([sys.executable, __file__, nm, *[a for a in sys.argv[1:] if a not in names]])
    sys.exit(0)
os.environ["CC4_B200_LIB"] = lib(names[0])
sys.path.insert(0, ROOT)
import numpy as np, torch
from cage_challenge_4_b200 import BatchedCC4, _native

n, warm, ticks = arg("--envs", 65536), arg("--warm", 150), arg("--ticks", 100)
env = BatchedCC4(n, device=0, seed=0)
env.reset()
g = torch.Generator(device="cuda").manual_seed(0)
def acts():
    return torch.stack([torch.randint(0, k, (n,), generator=g, device="cuda") for k in (82, 82, 82, 82, 242)], 1).to(torch.int32).contiguous()
for _ in range(warm):
    env.step_compact(acts())
buf = np.zeros(16, dtype=np.uint64)
_native.check(env._L.cc4_phase_clocks(env._h, buf.ctypes.data, 1), "cc4_phase_clocks")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
pool = [acts() for _ in range(ticks)]
e0.record()
for a in pool:
    env.step_compact(a)
e1.record()
torch.cuda.synchronize()
_native.check(env._L.cc4_phase_clocks(env._h, buf.ctypes.data, 1), "cc4_phase_clocks")
blocks = float(buf[8])
labels = ["load+setup", "policies", "blue", "green", "red", "reassign", "end turn", "finish+obs+store"]
tot = float(buf[9]) / blocks
print(f"[{names[0]}] {n} episodes, {ticks} ticks, {e0.elapsed_time(e1) / ticks:.4f} ms per tick; {blocks / ticks:.0f} blocks per tick, block wall {tot:.0f} cycles")
for i, l in enumerate(labels):
    w = float(buf[i]) / blocks
    if w > 0:
        print(f"  {l:>18} {w:9.0f} cycles {100 * w / tot:5.1f} %   (with CC4_FLOW a mark covers every stage up to the next barrier)")
assert not env.error_flags().any()
