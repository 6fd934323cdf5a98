#!/usr/bin/env bash
# round 2, GPU call B: half-tile GEMM (2 CTAs/SM) + cluster substitution: correctness, then the bench line and small-n sweep
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $O/b_tests_old.log 2>&1
echo "old tests rc=$?" >> $O/b_tests_old.log
timeout 900 python -m pytest tests/test_gpu_api_r02.py tests/test_gpu_baseline_workloads.py -m gpu -q > $O/b_tests_new.log 2>&1
echo "new tests rc=$?" >> $O/b_tests_new.log
timeout 600 python bench.py --cpu-seconds 0 > $O/b_bench.json 2> $O/b_bench.err
timeout 300 python bench.py --trials 13 --cpu-seconds 0 --steps 10 > $O/b_bench_cp13.json 2> $O/b_bench_cp13.err
GPRB200_REUSE=0 timeout 300 python tools/latency.py > $O/b_latency.json 2> $O/b_latency.err
timeout 600 python tools/sweep.py --dims 26 --nmax 1024 --out $O/b_sweep_small.json > $O/b_sweep_small.log 2>&1
GPRB200_LIB=$PWD/gpr.jl_b200/libgprb200_tl.so timeout 300 python bench.py --trials 25 --cpu-seconds 0 --steps 1 --warmup 1 --no-predict > $O/b_tl.json 2> $O/b_tl.log
tail -n 3 $O/b_tests_old.log $O/b_tests_new.log
head -c 600 $O/b_bench.json
