#!/usr/bin/env bash
# round 2, GPU call D: tensor-map TMA producer in the half-tile GEMM
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $O/d_tests_old.log 2>&1
echo "old tests rc=$?" >> $O/d_tests_old.log
timeout 900 python -m pytest tests/test_gpu_api_r02.py tests/test_gpu_baseline_workloads.py -m gpu -q > $O/d_tests_new.log 2>&1
echo "new tests rc=$?" >> $O/d_tests_new.log
timeout 600 python bench.py --cpu-seconds 0 > $O/d_bench.json 2> $O/d_bench.err
timeout 300 python bench.py --trials 13 --cpu-seconds 0 --steps 10 --no-predict > $O/d_bench_cp13.json 2> $O/d_bench_cp13.err
GPRB200_REUSE=0 timeout 300 python tools/latency.py > $O/d_latency.json 2> $O/d_latency.err
timeout 600 python tools/sweep.py --dims 26 --nmax 1024 --out $O/d_sweep_small.json > $O/d_sweep_small.log 2>&1
GPRB200_LIB=$PWD/gpr.jl_b200/libgprb200_tl.so timeout 300 python bench.py --trials 25 --cpu-seconds 0 --steps 1 --warmup 1 --no-predict > $O/d_tl.json 2> $O/d_tl.log
tail -n 3 $O/d_tests_old.log $O/d_tests_new.log
head -c 300 $O/d_bench.json
