#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 300 python bench.py --cpu-seconds 0 --no-predict > $O/h_bench_new.json 2> $O/h_bench_new.err
timeout 300 python bench.py --trials 13 --steps 10 --cpu-seconds 0 --no-predict > $O/h_bench_13.json 2>> $O/h_bench_new.err
