#!/usr/bin/env bash
# round 2, multi-GPU call: weak and strong scaling of the CP configuration, strong scaling of FB; N = $1
set -u
N=${1:-8}
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
run() { # tag, extra args
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) bench.py --gpus $N --steps 5 --warmup 3 --cpu-seconds 0 $2 > $O/k${N}_$1.json 2> $O/k${N}_$1.err
  tail -n 1 $O/k${N}_$1.json | cut -c1-200
}
run weak ""
run strong "--scaling strong"
run fb_strong "--scaling strong --system FB --no-predict"
