#!/usr/bin/env bash
# Extra ncu captures (run under gpurun, ONE gpu): a CHOL_DIAG + CHOL_COL launch pair of the Cholesky stage and one
# FWD_ROW launch of the predictive variance (m = 100 test columns, compact column layout).  Usage: tools/profile_extra.sh <tag>
set -uo pipefail
TAG=${1:-r01}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --trials 12 --cpu-seconds 0 --no-predict"
$CMD > gpurun_out/${TAG}_plain4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_tile_gemm -s 20 -c 2 -f -o gpurun_out/${TAG}_k6 $CMD > gpurun_out/${TAG}_ncu_k6.log 2>&1
echo "capture chol rc=$?"
PCMD="python tools/predict_bench.py --trials 40 --ms 100"
$PCMD > gpurun_out/${TAG}_plain5.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_tile_gemm -s 43 -c 1 -f -o gpurun_out/${TAG}_k7 $PCMD > gpurun_out/${TAG}_ncu_k7.log 2>&1
echo "capture fwd_row rc=$?"
