#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 60 ./tools/fp64_micro > $O/c_micro.json 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $O/c_tests_old.log 2>&1
echo "old tests rc=$?" >> $O/c_tests_old.log
timeout 900 python -m pytest tests/test_gpu_api_r02.py tests/test_gpu_baseline_workloads.py -m gpu -q > $O/c_tests_new.log 2>&1
echo "new tests rc=$?" >> $O/c_tests_new.log
GPRB200_SOLVE_CLUSTER_BELOW=0 timeout 600 python bench.py --cpu-seconds 0 --no-predict > $O/c_bench_nocluster.json 2> $O/c_bench.err
timeout 600 python bench.py --cpu-seconds 0 --no-predict > $O/c_bench_cluster.json 2>> $O/c_bench.err
cat $O/c_micro.json
tail -n 3 $O/c_tests_old.log $O/c_tests_new.log
