#!/usr/bin/env bash
# ncu evidence for one round (run under gpurun, ONE gpu).  Every ncu pass is preceded by the same command without ncu.
#   1. launch list (gpu__time_duration of every kernel of one small bench run)      -> gpurun_out/<tag>_launches.csv
#   2. full capture of three k_tile_gemm launches (TRTRI row 14, 15, LAUUM; 12 GPs) -> gpurun_out/<tag>_gemm.ncu-rep
#   3. full capture of one launch of each other kernel                              -> gpurun_out/<tag>_k<i>.ncu-rep
# Usage: tools/profile.sh <tag>
set -uo pipefail
TAG=${1:-r02}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --trials 12 --cpu-seconds 0"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 420 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_list.log 2>&1
echo "launch list rc=$?"
FULL="python bench.py --steps 1 --warmup 3 --cpu-seconds 0 --no-predict"   # 1b. the same list at the full bench configuration
$FULL > gpurun_out/${TAG}_full_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_B400.csv $FULL > gpurun_out/${TAG}_full_ncu.log 2>&1
echo "full launch list rc=$?"
$CMD > gpurun_out/${TAG}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_tile_gemm -s 43 -c 3 -f -o gpurun_out/${TAG}_gemm $CMD > gpurun_out/${TAG}_ncu_gemm.log 2>&1
echo "gemm capture rc=$?"
i=0
for K in k_grad_tiles k_assemble k_diag_factor k_solve k_predict_cross k_predict_finish; do
  $CMD > gpurun_out/${TAG}_plain3.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -f -o gpurun_out/${TAG}_k$i $CMD > gpurun_out/${TAG}_ncu_k$i.log 2>&1
  echo "capture $K rc=$?"
  i=$((i+1))
done
tail -1 gpurun_out/${TAG}_plain.log | cut -c1-300
