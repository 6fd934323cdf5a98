#!/usr/bin/env bash
# strong scaling with the GP-level split (equal GP ranges per rank); N = $1
set -u
N=${1:-8}
cd "$(dirname "$0")/.."
O=gpurun_out
run() { # tag, extra args
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) bench.py --gpus $N --steps 5 --warmup 3 --cpu-seconds 0 $2 > $O/u${N}_$1.json 2> $O/u${N}_$1.err
  tail -n 1 $O/u${N}_$1.json | cut -c1-200
}
run strong_gp "--scaling strong --split gp --no-predict"
run fb_strong_gp "--scaling strong --split gp --system FB --no-predict"
