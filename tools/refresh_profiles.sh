#!/usr/bin/env bash
# Refresh every measured artefact of a round (run under gpurun, ONE gpu): bench line, ncu launch lists + captures,
# DRAM traffic of the tile GEMM, per-launch GEMM breakdown, FB configuration.  Usage: tools/refresh_profiles.sh <tag>
set -uo pipefail
TAG=${1:-r01z}
mkdir -p gpurun_out
bash tools/profile.sh $TAG 2>&1 | tail -12
bash tools/traffic.sh $TAG 2>&1 | tail -2
python tools/gemm_breakdown.py --out gpurun_out/${TAG}_gemm_breakdown.json > gpurun_out/${TAG}_breakdown.log 2>&1; tail -7 gpurun_out/${TAG}_breakdown.log
python bench.py --system FB --cpu-seconds 0 > gpurun_out/${TAG}_bench_FB.log 2>&1; tail -1 gpurun_out/${TAG}_bench_FB.log | cut -c1-160
python bench.py > gpurun_out/${TAG}_bench_B400.log 2>&1; tail -1 gpurun_out/${TAG}_bench_B400.log | cut -c1-300
