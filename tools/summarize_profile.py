#!/usr/bin/env python
"""Write profiles/<tag>_summary.md from the artefacts of tools/profile.sh (launch list csv + .ncu-rep captures).
Usage: tools/summarize_profile.py <tag>   (reads gpurun_out/<tag>_*, copies the launch list into profiles/)"""
import collections
import csv
import glob
import io
import os
import shutil
import subprocess
import sys

tag = sys.argv[1]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
go, pr = os.path.join(root, "gpurun_out"), os.path.join(root, "profiles")
lines = [f"# {tag}: ncu evidence (tools/profile.sh {tag}; bench.py --steps 1 --warmup 1 --trials 12 => B=48 GPs, n=2000, d=26, one B200)", ""]

rows = [r for r in csv.reader(open(os.path.join(go, f"{tag}_launches.csv"))) if len(r) > 5]
hdr = next(r for r in rows if r[0] == "ID")
kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows:
    if r[0] == "ID":
        continue
    v = float(r[mv].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(r[mu], 1e-6)
    name = r[kn].split("(")[0].replace("void ", "")
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(v[1] for v in agg.values())
lines += ["## Launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`; cold-cache and serialised: compare shares)", "",
          "| kernel | launches | total ms | share |", "|---|---|---|---|"]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append(f"| {k} | {v[0]} | {v[1]:.3f} | {v[1] / tot:.3f} |")
shutil.copy(os.path.join(go, f"{tag}_launches.csv"), os.path.join(pr, f"{tag}_launches.csv"))

full = os.path.join(go, f"{tag}_launches_B400.csv")
if os.path.exists(full):  # launch list of the FULL bench configuration (tools/profile.sh step 1b)
    rows = [r for r in csv.reader(open(full)) if len(r) > 5]
    agg = collections.OrderedDict()
    for r in rows:
        if r[0] == "ID":
            continue
        v = float(r[mv].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(r[mu], 1e-6)
        name = r[kn].split("(")[0].replace("void ", "").replace("gprb::", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    lines += ["", "## Launch list at the full bench configuration (`bench.py --steps 1 --warmup 3 --no-predict`: B=400 GPs, n=2000, d=26)", "",
              "The share of `k_tile_gemm` here is to be compared with the CUDA-event stage times `bench.py` reports "
              "(`roofline.stage_ms`: gemm / total).", "", "| kernel | launches | total ms | share |", "|---|---|---|---|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if v[1] / tot >= 0.0005:
            lines.append(f"| {k} | {v[0]} | {v[1]:.3f} | {v[1] / tot:.3f} |")
    shutil.copy(full, os.path.join(pr, f"{tag}_launches_B400.csv"))

WANT = [("gpu__time_duration.sum", "duration"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__registers_per_thread", "regs/thread"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 (DFMA) pipe active %"),
        ("SM_C.TriageCompute.smsp__pipe_tensor_subpipe_dmma_cycles_active.avg", "DMMA cycles active (avg/SMSP)"),
        ("sm__cycles_active.avg", "SM cycles active (avg)"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts")]
STALL = "smsp__average_warps_issue_stalled_"
for rep in sorted(glob.glob(os.path.join(go, f"{tag}_gemm.ncu-rep")) + glob.glob(os.path.join(go, f"{tag}_k*.ncu-rep"))):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(io.StringIO(out)))
    if len(rr) < 3:
        continue
    h, u = rr[0], rr[1]
    legend = {"gemm": "tile GEMM: TRTRI_ROW rows 14 and 15, then the LAUUM launch (12 GPs)", "k0": "gradient tiles", "k1": "covariance assembly",
              "k2": "diagonal-block factor", "k3": "substitution + mll", "k4": "predict: cross-covariances", "k5": "predict: finishing reduction",
              "k6": "tile GEMM, Cholesky stage: CHOL_DIAG and CHOL_COL of block column 10 (tools/profile_extra.sh)",
              "k7": "tile GEMM, predictive variance: one FWD_ROW launch, m = 100 test columns, 40 GPs of one stream group (tools/profile_extra.sh)"}
    key = os.path.basename(rep)[len(tag) + 1:-len(".ncu-rep")]
    lines += ["", f"## `ncu --set full` : {os.path.basename(rep)}" + (f" - {legend[key]}" if key in legend else ""), ""]
    for r in rr[2:]:
        name = r[h.index("Kernel Name")]
        lines.append(f"**{name}**")
        lines.append("")
        for key, label in WANT:
            if key in h:
                i = h.index(key)
                lines.append(f"- {label}: {r[i]} {u[i]}")
        if "SM_C.TriageCompute.smsp__pipe_tensor_subpipe_dmma_cycles_active.avg" in h and "sm__cycles_active.avg" in h:
            try:
                d = float(r[h.index("SM_C.TriageCompute.smsp__pipe_tensor_subpipe_dmma_cycles_active.avg")].replace(",", ""))
                a = float(r[h.index("sm__cycles_active.avg")].replace(",", ""))
                if d > 0:
                    lines.append(f"- DMMA pipe busy while the SM is active: {d / a:.3f}")
            except ValueError:
                pass
        st = [(float(r[i]), x[len(STALL):-len("_per_issue_active.ratio")]) for i, x in enumerate(h)
              if x.startswith(STALL) and x.endswith("_per_issue_active.ratio") and r[i] not in ("", "n/a")]
        lines.append("- top stalls per issued instruction: " + ", ".join(f"{n} {v:.2f}" for v, n in sorted(st, reverse=True)[:5]))
        lines.append("")
open(os.path.join(pr, f"{tag}_summary.md"), "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:40]))
