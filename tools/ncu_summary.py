#!/usr/bin/env python
"""Text summary of an `ncu --set full` capture (the metrics bench.py's roofline quotes):  python tools/ncu_summary.py rep.ncu-rep > profiles/rNN_kstep_ncu_full_summary.txt"""
import csv, io, subprocess, sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "lts__t_sector_hit_rate.pct", "sm__inst_executed.avg.per_cycle_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
        "launch__registers_per_thread", "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
        "sm__icc_request_hit_rate.pct", "sm__icc_requests.sum", "smsp__average_warp_latency_per_inst_issued.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "dram__throughput.avg.pct_of_peak_sustained_elapsed"]
txt = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units = rows[0], rows[1]
for k, r in enumerate(rows[2:]):
    d = dict(zip(hdr, r))
    print(f"-- captured launch {k} {d.get('Kernel Name')}")
    for w in WANT:
        if w in hdr:
            print(w, units[hdr.index(w)], d[w])
    st = {h.replace("smsp__pcsamp_warps_issue_stalled_", ""): float(d[h]) for h in hdr
          if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued") and d[h] not in ("", "n/a")}
    tot = sum(st.values()) or 1.0
    print("stall samples: " + ", ".join(f"{k2} {100 * v / tot:.1f}%" for k2, v in sorted(st.items(), key=lambda kv: -kv[1])[:10]))
