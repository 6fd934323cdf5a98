#!/usr/bin/env bash
# final verification of the committed tree: full GPU suite, smoke(), the bench lines (default with the CPU arm, FB, the strong-split
# sizes), the prediction sweep and the ncu capture of the (changed) cross-covariance kernel
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > $O/final_gputests.log 2>&1; echo "gpu tests rc=$?" >> $O/final_gputests.log; tail -2 $O/final_gputests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/final_smoke.log 2>&1; echo "smoke rc=$?" >> $O/final_smoke.log; tail -2 $O/final_smoke.log
timeout 600 python bench.py > $O/final_bench.json 2> $O/final_bench.err; tail -1 $O/final_bench.json | cut -c1-200
timeout 600 python bench.py --system FB --cpu-seconds 0 > $O/final_bench_FB_B1200.json 2>> $O/final_bench.err
timeout 300 python bench.py --trials 13 --steps 10 --cpu-seconds 0 > $O/final_bench_CP_B52.json 2>> $O/final_bench.err
timeout 300 python bench.py --system FB --trials 13 --steps 5 --cpu-seconds 0 > $O/final_bench_FB_B156.json 2>> $O/final_bench.err
timeout 600 python tools/predict_bench.py --out $O/final_predict.json > $O/final_predict.log 2>&1
GPRB200_REUSE=0 timeout 300 python tools/latency.py $O/final_latency.json > $O/final_latency.log 2>&1
CMD="python bench.py --steps 1 --warmup 1 --trials 12 --cpu-seconds 0"
$CMD > $O/final_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_predict_cross -s 1 -c 1 -f -o $O/r02_k4 $CMD > $O/final_ncu_k4.log 2>&1
echo "capture k_predict_cross rc=$?"
