#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api_r02.py -m gpu -x -q > $O/m_tests.log 2>&1
echo "tests rc=$?" >> $O/m_tests.log
timeout 600 python bench.py --cpu-seconds 0 > $O/m_bench.json 2> $O/m_bench.err
timeout 300 python bench.py --trials 13 --cpu-seconds 0 --steps 10 --no-predict > $O/m_bench_cp13.json 2> $O/m_bench_cp13.err
GPRB200_REUSE=0 timeout 300 python tools/latency.py > $O/m_latency.json 2> $O/m_latency.err
timeout 600 python tools/sweep.py --dims 26 --nmax 1024 --out $O/m_sweep_small.json > $O/m_sweep_small.log 2>&1
GPRB200_LIB=$PWD/gpr.jl_b200/libgprb200_tl.so timeout 300 python bench.py --trials 25 --cpu-seconds 0 --steps 1 --warmup 1 --no-predict > $O/m_tl.json 2> $O/m_tl.log
tail -n 3 $O/m_tests.log
