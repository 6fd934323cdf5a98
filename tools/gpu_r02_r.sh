#!/usr/bin/env bash
# second half of the round-2 refresh: the n >= 2048 sweep points at the 48 GB batch sizes, smoke(), the reference arm
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 1200 python tools/sweep.py --nmin 2048 --gb 48 --out $O/r02_sweep_big.json > $O/r02_sweep_big.log 2>&1; tail -2 $O/r02_sweep_big.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_smoke.log 2>&1; echo "smoke rc=$?" >> $O/r02_smoke.log; tail -2 $O/r02_smoke.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/r02_bench_reference_arm.json 2> $O/r02_bench_ref.err; tail -1 $O/r02_bench_reference_arm.json | cut -c1-300
