#!/usr/bin/env bash
# round 2, GPU call A: new parity / API tests, the full GPU suite, the default bench line, per-GPU rates of the strong split
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/a_smi.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_api_r02.py tests/test_gpu_baseline_workloads.py -m gpu -x -q > $O/a_tests_new.log 2>&1
echo "new tests rc=$?" >> $O/a_tests_new.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $O/a_tests_old.log 2>&1
echo "old tests rc=$?" >> $O/a_tests_old.log
timeout 600 python bench.py > $O/a_bench.json 2> $O/a_bench.err
timeout 300 python bench.py --trials 13 --cpu-seconds 0 --steps 10 > $O/a_bench_cp13.json 2> $O/a_bench_cp13.err
timeout 300 python bench.py --trials 12 --cpu-seconds 0 --steps 10 > $O/a_bench_cp12.json 2>> $O/a_bench_cp13.err
timeout 300 python bench.py --system FB --trials 13 --cpu-seconds 0 --steps 5 > $O/a_bench_fb13.json 2> $O/a_bench_fb13.err
GPRB200_REUSE=0 timeout 300 python tools/latency.py > $O/a_latency.json 2> $O/a_latency.err
tail -3 $O/a_tests_new.log $O/a_tests_old.log
head -c 1500 $O/a_bench.json
