#!/usr/bin/env bash
# Full ncu capture of ONE launch of each named kernel (run under gpurun, one GPU).
# Usage: tools/profile_kernel.sh <tag> <kernel-regex> [<kernel-regex> ...]  -> gpurun_out/<tag>_<n>.ncu-rep
set -uo pipefail
TAG=$1; shift
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --trials 12 --no-predict --cpu-seconds 0"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
n=0
for K in "$@"; do
  ncu --set full --clock-control none --import-source on -k regex:"$K" -s 2 -c 1 -f -o gpurun_out/${TAG}_$n $CMD > gpurun_out/${TAG}_ncu_$n.log 2>&1
  echo "capture $K rc=$?"
  n=$((n+1))
done
