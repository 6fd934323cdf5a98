#!/usr/bin/env python
"""Key metrics of every launch in an ncu report.  Usage: tools/ncu_metrics.py <report.ncu-rep>"""
import csv
import io
import subprocess
import sys

WANT = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "sm__cycles_active.avg", "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "SM_C.TriageCompute.smsp__pipe_tensor_subpipe_dmma_cycles_active.avg", "smsp__issue_active.avg.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum"]
STALL = "smsp__average_warps_issue_stalled_"
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print("=" * 100)
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print(f"  {w:75s} {r[i]} {units[i]}")
    st = [(float(r[i]), h[len(STALL):-len('_per_issue_active.ratio')]) for i, h in enumerate(hdr)
          if h.startswith(STALL) and h.endswith("_per_issue_active.ratio") and r[i] not in ("", "n/a")]
    print("  stalls per issue:", ", ".join(f"{n}={v:.2f}" for v, n in sorted(st, reverse=True)[:7]))
