#!/usr/bin/env python
"""Parse an ncu --csv metric log of k_tile_gemm launches; keep the launches of the last (profiled) evaluation."""
import csv
import json
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 8]
hdr = next(r for r in rows if r[0] == "ID")
iid, name, met, unit, val = (hdr.index(k) for k in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value"))
launches = {}
for r in rows:
    if r[0] == "ID":
        continue
    v = float(r[val].replace(",", ""))
    u = r[unit]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1)
    launches.setdefault(int(r[iid]), {})[r[met]] = v * scale
NL = int(sys.argv[3]) if len(sys.argv) > 3 else 46  # tile-GEMM launches of one evaluation at n = 2000: 15 CHOL_DIAG + 15 CHOL_COL + 15 TRTRI_ROW + LAUUM
ids = sorted(launches)[-NL:]
rd = sum(launches[i]["dram__bytes_read.sum"] for i in ids)
wr = sum(launches[i]["dram__bytes_write.sum"] for i in ids)
ms = sum(launches[i]["gpu__time_duration.sum"] for i in ids)
out = {"kernel": "k_tile_gemm", "launches": len(ids), "total_launches_seen": len(launches),
       "dram_read_bytes_per_launch": rd / len(ids), "dram_write_bytes_per_launch": wr / len(ids),
       "dram_bytes_per_launch": (rd + wr) / len(ids), "ncu_ms_per_launch": ms / len(ids),
       "note": "the launches of one profiled evaluation (single stream, all GPs of the batch per launch), ncu cold-cache"}
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps(out))
