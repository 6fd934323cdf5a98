#!/usr/bin/env bash
# final verification of the committed tree: full GPU suite, smoke(), the default bench line (with the CPU arm), --impl reference
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > $O/final_gputests.log 2>&1; echo "gpu tests rc=$?" >> $O/final_gputests.log; tail -2 $O/final_gputests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/final_smoke.log 2>&1; echo "smoke rc=$?" >> $O/final_smoke.log; tail -2 $O/final_smoke.log
timeout 600 python bench.py > $O/final_bench.json 2> $O/final_bench.err; tail -1 $O/final_bench.json | cut -c1-200
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/final_bench_ref.json 2> $O/final_bench_ref.err; tail -1 $O/final_bench_ref.json | cut -c1-300
timeout 300 python bench.py --trials 13 --steps 10 --cpu-seconds 0 > $O/final_bench_CP_B52.json 2>> $O/final_bench.err
