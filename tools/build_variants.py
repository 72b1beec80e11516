#!/usr/bin/env python
"""Build tuning variants of libcc4_b200.so (same C-ABI, different -D knobs) into cage_challenge_4_b200/variants/ and,
with --bench, time each one (run on the GPU box: `python tools/build_variants.py --bench`).

    python tools/build_variants.py name1="-DCC4_WARPS_PER_BLOCK=4" name2="-DCC4_SYNC_MASK=0x55" ...
"""
import json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "cage_challenge_4_b200")
VDIR = os.path.join(PKG, "variants")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared", "-Xcompiler", "-fPIC"]


def build(name, defs):
    os.makedirs(VDIR, exist_ok=True)
    out = os.path.join(VDIR, f"libcc4_b200_{name}.so")
    subprocess.check_call(["nvcc", *FLAGS, *defs.split(), "-o", out, os.path.join(PKG, "csrc", "cc4_engine.cu")])
    return out


def bench(path, extra):
    env = dict(os.environ, CC4_B200_LIB=path)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--no-cpu", "--steps", "120", "--warmup", "5", "--e2e-steps", "40", *extra],
                       env=env, capture_output=True, text=True)
    try:
        j = json.loads(r.stdout.strip().split("\n")[-1])
        return {"value": j["value"], "ms_per_step": j["ms_per_step"], "kernel_ms": j["roofline"]["kernel_ms"], "e2e": j["e2e"]["value"], "err": j["error_episodes"], "score": j["score"]["mean_return"]}
    except Exception as e:  # noqa: BLE001
        return {"error": str(e), "stderr": r.stderr[-400:]}


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    if "--bench" in sys.argv:
        extra = []
        for f in sorted(os.listdir(VDIR)):
            if f.endswith(".so"):
                print(f, json.dumps(bench(os.path.join(VDIR, f), extra)), flush=True)
    else:
        for a in args:
            name, defs = a.split("=", 1)
            print(build(name, defs))
