#!/usr/bin/env bash
# Refresh every measured artefact of a round (run under gpurun, ONE gpu): GPU test log, bench lines (CP, FB, strong-split sizes),
# ncu launch lists + captures, DRAM traffic of the tile GEMM, per-launch GEMM breakdown and tile timeline, sweep, latency,
# rollout, optimiser, prediction.  Usage: tools/refresh_profiles.sh <tag>
set -uo pipefail
TAG=${1:-r02}
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/${TAG}_gputests.log 2>&1; echo "gpu tests rc=$?" >> $O/${TAG}_gputests.log; tail -2 $O/${TAG}_gputests.log
python bench.py > $O/${TAG}_bench_B400.json 2> $O/${TAG}_bench_B400.err; tail -1 $O/${TAG}_bench_B400.json | cut -c1-200
python bench.py --system FB --cpu-seconds 0 > $O/${TAG}_bench_FB_B1200.json 2> $O/${TAG}_bench_FB.err; tail -1 $O/${TAG}_bench_FB_B1200.json | cut -c1-160
python bench.py --trials 13 --steps 10 --cpu-seconds 0 > $O/${TAG}_bench_CP_B52.json 2>> $O/${TAG}_bench_FB.err
python bench.py --system FB --trials 13 --steps 5 --cpu-seconds 0 > $O/${TAG}_bench_FB_B156.json 2>> $O/${TAG}_bench_FB.err
bash tools/profile.sh $TAG 2>&1 | tail -12
bash tools/traffic.sh $TAG 2>&1 | tail -2
python tools/gemm_breakdown.py --out $O/${TAG}_gemm_breakdown.json > $O/${TAG}_breakdown.log 2>&1; tail -7 $O/${TAG}_breakdown.log
GPRB200_LIB=$PWD/gpr.jl_b200/libgprb200_tl.so python bench.py --trials 25 --cpu-seconds 0 --steps 1 --warmup 1 --no-predict > /dev/null 2> $O/${TAG}_gemm_tile_timeline.log
GPRB200_REUSE=0 python tools/latency.py $O/${TAG}_latency.json > $O/${TAG}_latency.log 2>&1
python tools/rollout_bench.py --newton 1 --out $O/${TAG}_rollout_FB.json > $O/${TAG}_rollout.log 2>&1
python tools/rollout_bench.py --system CP --newton 1 --out $O/${TAG}_rollout_CP.json >> $O/${TAG}_rollout.log 2>&1
python tools/optimize_bench.py --out $O/${TAG}_optimize.json > $O/${TAG}_optimize.log 2>&1
python tools/predict_bench.py --out $O/${TAG}_predict.json > $O/${TAG}_predict.log 2>&1
python tools/sweep.py --out $O/${TAG}_sweep.json > $O/${TAG}_sweep.log 2>&1; tail -3 $O/${TAG}_sweep.log | cut -c1-160
