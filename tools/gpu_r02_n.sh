#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api_r02.py -m gpu -x -q > $O/n_tests.log 2>&1
echo "tests rc=$?" >> $O/n_tests.log
timeout 300 python bench.py --cpu-seconds 0 --no-predict > $O/n_bench.json 2> $O/n_bench.err
timeout 300 python bench.py --trials 13 --cpu-seconds 0 --steps 10 --no-predict > $O/n_bench_cp13.json 2>> $O/n_bench.err
timeout 300 python bench.py --trials 12 --cpu-seconds 0 --steps 10 --no-predict > $O/n_bench_cp12.json 2>> $O/n_bench.err
timeout 300 python bench.py --system FB --trials 13 --cpu-seconds 0 --steps 5 --no-predict > $O/n_bench_fb13.json 2>> $O/n_bench.err
timeout 600 python tools/sweep.py --dims 26 --nmax 1024 --out $O/n_sweep_small.json > $O/n_sweep_small.log 2>&1
tail -n 3 $O/n_tests.log
