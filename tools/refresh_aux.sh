#!/usr/bin/env bash
# Secondary measurements (run under gpurun, one GPU): FB configuration, small-batch latency, batched optimiser, predict sweep.
set -uo pipefail
mkdir -p gpurun_out
python bench.py --system FB --cpu-seconds 0 > gpurun_out/r01z_bench_FB.log 2>&1; tail -1 gpurun_out/r01z_bench_FB.log | cut -c1-200
python tools/latency.py gpurun_out/r01z_latency.json 2>&1 | tail -3
python tools/optimize_bench.py --out gpurun_out/r01z_optimize.json 2>&1 | tail -2
python tools/predict_bench.py --out gpurun_out/r01z_predict.json 2>&1 | tail -5
