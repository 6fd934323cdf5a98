#!/usr/bin/env bash
# round 2, GPU call F: restored substitution order + column-split cluster backward; rollout / optimise benches; stream-count probes
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api_r02.py -m gpu -x -q > $O/f_tests.log 2>&1
echo "tests rc=$?" >> $O/f_tests.log
timeout 600 python bench.py --cpu-seconds 0 > $O/f_bench.json 2> $O/f_bench.err
timeout 300 python bench.py --trials 13 --cpu-seconds 0 --steps 10 --no-predict > $O/f_bench_cp13.json 2> $O/f_bench_cp13.err
GPRB200_STREAMS=2 timeout 300 python bench.py --trials 13 --cpu-seconds 0 --steps 10 --no-predict > $O/f_bench_cp13_s2.json 2>> $O/f_bench_cp13.err
GPRB200_RL_MAX=64 timeout 300 python bench.py --trials 13 --cpu-seconds 0 --steps 10 --no-predict > $O/f_bench_cp13_rl.json 2>> $O/f_bench_cp13.err
GPRB200_STREAMS=8 timeout 300 python bench.py --cpu-seconds 0 --steps 3 > $O/f_bench_s8.json 2> $O/f_bench_s8.err
GPRB200_REUSE=0 timeout 300 python tools/latency.py > $O/f_latency.json 2> $O/f_latency.err
timeout 600 python tools/rollout_bench.py --newton 1 --out $O/f_rollout_fb.json > $O/f_rollout_fb.log 2>&1
timeout 600 python tools/rollout_bench.py --system CP --newton 1 --out $O/f_rollout_cp.json > $O/f_rollout_cp.log 2>&1
timeout 600 python tools/optimize_bench.py > $O/f_optimize.log 2>&1
tail -n 3 $O/f_tests.log
head -c 300 $O/f_bench.json
