#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api_r02.py tests/test_gpu_baseline_workloads.py -m gpu -x -q > $O/p_tests.log 2>&1
echo "tests rc=$?" >> $O/p_tests.log
timeout 600 python bench.py --cpu-seconds 0 --no-predict > $O/p_bench.json 2> $O/p_bench.err
timeout 300 python bench.py --trials 13 --cpu-seconds 0 --steps 10 --no-predict > $O/p_bench_cp13.json 2> $O/p_bench_cp13.err
GPRB200_REUSE=0 timeout 300 python tools/latency.py > $O/p_latency.json 2> $O/p_latency.err
timeout 500 python tools/small_n_breakdown.py > $O/p_small_bd.log 2>&1
tail -n 3 $O/p_tests.log
