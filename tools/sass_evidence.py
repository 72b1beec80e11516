#!/usr/bin/env python
"""Instruction counts per kernel of the shipped libcc4_b200.so (cuobjdump -sass): python tools/sass_evidence.py > profiles/sass_rNN.txt"""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "cage_challenge_4_b200", "libcc4_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "--dump-resource-usage", lib], capture_output=True, text=True).stdout
usage = {}
for m in re.finditer(r"Function (\S+):\s*\n\s*(.*)", res):
    usage[m.group(1)] = m.group(2).strip()
KEYS = ["UBLKCP", "SYNCS", "BAR", "VOTE", "MATCH", "SHFL", "REDUX", "ATOMS", "LDL", "STL", "CALL", "LDG", "STG", "LDS", "STS", "LDC", "IMAD", "LOP3", "UTCMMA", "LDTM"]
cnt, n = collections.defaultdict(collections.Counter), collections.Counter()
fn = None
for l in sass.split("\n"):
    m = re.search(r"Function : (\S+)", l)
    if m:
        fn = m.group(1); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", l)
    if m and fn:
        n[fn] += 1
        op = m.group(1)
        for k in KEYS:
            if op == k or op.startswith(k + "."):
                cnt[fn][k] += 1
print("SASS evidence of the shipped cage_challenge_4_b200/libcc4_b200.so (cuobjdump -sass, sm_100a, CUDA 12.9; nvcc -O3 -lineinfo).")
print("Per kernel: code size, resources, and the counts of the instructions that show the bulk-copy engine (UBLKCP = cp.async.bulk, 1-D TMA), the")
print("mbarrier (SYNCS), block barriers (BAR), warp votes / matches / shuffles / reductions (the warp-uniform scans), local-memory traffic")
print("(LDL / STL: register spills), and calls (CALL: the out-of-line Philox block, refills, tape / variant forms).  Namespace cc4p = the production")
print("kernels (no tape paths), cc4 = the same with the tape paths + the reset kernels.  No dense contraction on this path: no UTCMMA / LDTM expected.\n")
for f in sorted(n, key=lambda k: (("cc4p" not in k), k)):
    print(f)
    print(f"  {n[f]} instructions = {n[f] * 16 / 1024:.1f} KB; {usage.get(f, '')}")
    print("  " + ", ".join(f"{k} {cnt[f][k]}" for k in KEYS) + "\n")
