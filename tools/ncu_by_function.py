#!/usr/bin/env python
"""Per-function / per-line digest of an `ncu --set full --import-source on` capture (needs -lineinfo).

    python tools/ncu_by_function.py gpurun_out/prof.ncu-rep [--lines 40]

Sums instructions executed and warp-stall samples of the FIRST captured kernel by the source function that
contains each line (a function = a line that starts `__device__`/`__global__` up to the next such line).
"""
import csv, io, re, subprocess, sys, collections

def main():
    rep = sys.argv[1]
    nlines = int(sys.argv[sys.argv.index("--lines") + 1]) if "--lines" in sys.argv else 30
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    files_seen, cur, hdr = set(), None, None
    per_line = collections.defaultdict(lambda: [0, 0, 0, collections.Counter()])   # inst, samples, thread-inst, stalls
    src_text, line_no = {}, None
    foot = collections.defaultdict(lambda: [0, 0])      # SASS instructions of the line: all, executed at least once
    for r in rows:
        if r and r[0] == "File Path":
            if files_seen and r[1] == first:   # the second captured kernel starts
                break
            cur = r[1]
            if not files_seen: first = cur
            files_seen.add(cur); continue
        if r and r[0] == "Line No":
            hdr = r; ix = {}
            for i, h in enumerate(hdr): ix.setdefault(h, i)
            stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
            continue
        if hdr is None or len(r) < len(hdr): continue
        if r[0] != "":
            line_no = int(r[0]); src_text[(cur, line_no)] = r[1]; continue
        if not r[ix["Instructions Executed"]].isdigit(): continue
        d = per_line[(cur, line_no)]
        foot[(cur, line_no)][0] += 1
        if int(r[ix["Instructions Executed"]]) > 0: foot[(cur, line_no)][1] += 1
        d[0] += int(r[ix["Instructions Executed"]]); d[1] += int(r[ix["# Samples"]]); d[2] += int(r[ix["Thread Instructions Executed"]])
        for s in stalls:
            v = int(r[ix[s]])
            if v: d[3][s] += v
    # function map
    func_of = {}
    for f in {k[0] for k in per_line}:
        try: lines = open(f).read().split("\n")
        except OSError: continue
        name = "?"
        for i, l in enumerate(lines, 1):
            m = re.match(r"\s*(?:template.*>)?\s*__(?:device|global)__.*?(\w+)\s*\(", l)
            if m and not l.startswith(" " * 8): name = m.group(1)
            func_of[(f, i)] = name
    tot_i = sum(d[0] for d in per_line.values()); tot_s = sum(d[1] for d in per_line.values())
    by_f = collections.defaultdict(lambda: [0, 0, 0, collections.Counter()])
    for k, d in per_line.items():
        fn = func_of.get(k, k[0].split("/")[-1])
        b = by_f[fn]; b[0] += d[0]; b[1] += d[1]; b[2] += d[2]; b[3].update(d[3])
    launches = None
    foot_f = collections.defaultdict(lambda: [0, 0])
    for k, (a, e) in foot.items():
        fn = func_of.get(k, k[0].split("/")[-1]); foot_f[fn][0] += a; foot_f[fn][1] += e
    fa, fe = sum(v[0] for v in foot_f.values()), sum(v[1] for v in foot_f.values())
    print(f"static code {fa} instructions = {fa * 16 / 1024:.1f} KB, executed at least once in this launch {fe} = {fe * 16 / 1024:.1f} KB (the SM's instruction cache has to supply these every block-tick)")
    print("executed footprint by function (KB): " + ", ".join(f"{fn} {v[1] * 16 / 1024:.1f}/{v[0] * 16 / 1024:.1f}" for fn, v in sorted(foot_f.items(), key=lambda kv: -kv[1][1])[:30]))
    print(f"total warp-instructions {tot_i}  samples {tot_s}")
    print(f"{'function':28s} {'inst%':>6s} {'samp%':>6s} {'lanes':>5s}  top stalls")
    for fn, b in sorted(by_f.items(), key=lambda kv: -kv[1][1])[:40]:
        st = ", ".join(f"{s[6:]} {100 * v / max(1, b[1]):.0f}%" for s, v in b[3].most_common(3))
        print(f"{fn:28s} {100 * b[0] / tot_i:6.1f} {100 * b[1] / tot_s:6.1f} {b[2] / max(1, b[0]):5.1f}  {st}")
    print("\nhot lines by samples")
    for k, d in sorted(per_line.items(), key=lambda kv: -kv[1][1])[:nlines]:
        st = ", ".join(f"{s[6:]} {100 * v / max(1, d[1]):.0f}%" for s, v in d[3].most_common(2))
        print(f"{k[0].split('/')[-1]}:{k[1]:4d} inst {100 * d[0] / tot_i:5.1f}% samp {100 * d[1] / tot_s:5.1f}% [{st}] {src_text.get(k, '').strip()[:110]}")

    if "--by-inst" in sys.argv:
        print("\nhot lines by instructions")
        for k, d in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:nlines]:
            print(f"{k[0].split('/')[-1]}:{k[1]:4d} inst {100 * d[0] / tot_i:5.1f}% samp {100 * d[1] / tot_s:5.1f}% lanes {d[2] / max(1, d[0]):4.1f} {src_text.get(k, '').strip()[:120]}")

main()
