#!/usr/bin/env bash
# ncu evidence for one round (run under gpurun, ONE gpu): launch list + full capture of the tile GEMM.
# Usage: tools/profile.sh <tag>   -> gpurun_out/<tag>_*.{csv,ncu-rep,log}
set -uo pipefail
TAG=${1:-r01}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --trials 12 --no-predict --cpu-seconds 0"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 330 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_list.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/${TAG}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_tile_gemm -s 44 -c 3 -f -o gpurun_out/${TAG}_gemm $CMD > gpurun_out/${TAG}_ncu_gemm.log 2>&1
echo "gemm capture rc=$?"
$CMD > gpurun_out/${TAG}_plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_grad_tiles|k_assemble|k_diag_factor|k_solve" -s 20 -c 4 -f -o gpurun_out/${TAG}_others $CMD > gpurun_out/${TAG}_ncu_others.log 2>&1
echo "others capture rc=$?"
tail -2 gpurun_out/${TAG}_plain.log
