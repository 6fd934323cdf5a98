#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api_r02.py -m gpu -x -q > $O/i_tests.log 2>&1
echo "tests rc=$?" >> $O/i_tests.log
timeout 300 python bench.py --cpu-seconds 0 --no-predict > $O/i_bench_la.json 2> $O/i_bench.err
GPRB200_LOOKAHEAD=0 timeout 300 python bench.py --cpu-seconds 0 --no-predict > $O/i_bench_nola.json 2>> $O/i_bench.err
timeout 300 python bench.py --trials 13 --steps 10 --cpu-seconds 0 --no-predict > $O/i_bench13_la.json 2>> $O/i_bench.err
GPRB200_LOOKAHEAD=0 timeout 300 python bench.py --trials 13 --steps 10 --cpu-seconds 0 --no-predict > $O/i_bench13_nola.json 2>> $O/i_bench.err
timeout 300 python bench.py --system FB --trials 13 --steps 5 --cpu-seconds 0 --no-predict > $O/i_benchfb13_la.json 2>> $O/i_bench.err
timeout 600 python tools/sweep.py --dims 26 --nmax 1024 --out $O/i_sweep_small.json > $O/i_sweep_small.log 2>&1
tail -n 3 $O/i_tests.log
