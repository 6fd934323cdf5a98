#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
for S in 4 6 8; do
GPRB200_STREAMS=$S timeout 300 python bench.py --trials 13 --steps 10 --cpu-seconds 0 --no-predict > $O/l_bench13_s$S.json 2>> $O/l_bench.err
done
GPRB200_STREAMS=8 GPRB200_GROUP_SKEW=0.5 timeout 300 python bench.py --trials 13 --steps 10 --cpu-seconds 0 --no-predict > $O/l_bench13_s8_sk.json 2>> $O/l_bench.err
GPRB200_STREAMS=8 GPRB200_SOLVE_CLUSTER_BELOW=0 timeout 300 python bench.py --trials 13 --steps 10 --cpu-seconds 0 --no-predict > $O/l_bench13_s8_nocl.json 2>> $O/l_bench.err
GPRB200_STREAMS=8 timeout 300 python bench.py --trials 25 --steps 10 --cpu-seconds 0 --no-predict > $O/l_bench25_s8.json 2>> $O/l_bench.err
GPRB200_STREAMS=4 timeout 300 python bench.py --trials 25 --steps 10 --cpu-seconds 0 --no-predict > $O/l_bench25_s4.json 2>> $O/l_bench.err
