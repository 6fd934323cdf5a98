#!/usr/bin/env bash
# round 2, final multi-GPU call: the multi-GPU tests, then weak and strong scaling of CP, strong scaling of FB; N = $1
set -u
N=${1:-8}
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_api_r02.py -m gpu -q -k "gather or two_contexts or multi" > $O/s${N}_tests.log 2>&1; echo "rc=$?" >> $O/s${N}_tests.log; tail -3 $O/s${N}_tests.log
run() { # tag, extra args
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) bench.py --gpus $N --steps 5 --warmup 3 --cpu-seconds 0 $2 > $O/s${N}_$1.json 2> $O/s${N}_$1.err
  tail -n 1 $O/s${N}_$1.json | cut -c1-200
}
run weak ""
run strong "--scaling strong"
run fb_strong "--scaling strong --system FB --no-predict"
