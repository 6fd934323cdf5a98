#!/usr/bin/env python
"""Top source lines by warp-stall samples from an ncu report (needs -lineinfo + --import-source on).
Usage: tools/ncu_lines.py <report.ncu-rep> [--kernel REGEX] [--skip N] [--top N]"""
import argparse
import csv
import io
import subprocess

ap = argparse.ArgumentParser()
ap.add_argument("rep")
ap.add_argument("--kernel", default=None)
ap.add_argument("--skip", type=int, default=0)
ap.add_argument("--top", type=int, default=30)
a = ap.parse_args()
cmd = ["ncu", "-i", a.rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--launch-skip", str(a.skip), "--launch-count", "1"]
if a.kernel:
    cmd += ["--kernel-name", f"regex:{a.kernel}"]
out = subprocess.run(cmd, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname, hdr, data = "", None, []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif len(r) > 4 and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) > 7 and r[0].isdigit():
        try:
            data.append((int(r[4]), int(r[7]), fname, int(r[0]), r[1].strip()[:100]))
        except ValueError:
            pass
tot = sum(d[0] for d in data) or 1
print(f"total stall samples {tot}, instructions executed {sum(d[1] for d in data)}")
for s, ie, f, ln, src in sorted(data, reverse=True)[: a.top]:
    print(f"{s:7d} {s / tot:6.3f} inst={ie:9d} {f}:{ln}: {src}")
