#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
for SK in 0 0.3 0.6; do
GPRB200_GROUP_SKEW=$SK timeout 300 python bench.py --cpu-seconds 0 --no-predict > $O/j_bench_sk$SK.json 2>> $O/j_bench.err
done
GPRB200_GROUP_SKEW=0.5 timeout 300 python bench.py --trials 13 --steps 10 --cpu-seconds 0 --no-predict > $O/j_bench13_sk0.5.json 2>> $O/j_bench.err
GPRB200_GROUP_SKEW=0.5 GPRB200_STREAMS=3 timeout 300 python bench.py --cpu-seconds 0 --no-predict > $O/j_bench_sk0.5_s3.json 2>> $O/j_bench.err
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $O/j_tests.log 2>&1
