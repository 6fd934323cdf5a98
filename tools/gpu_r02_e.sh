#!/usr/bin/env bash
# round 2, GPU call E (2 GPUs): in-library NCCL gather (single-process multi-device + torchrun ranks), strong / weak bench at N=2
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
nvidia-smi -L > $O/e_smi.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_api_r02.py -m gpu -q -k "gather or two_contexts or cluster" > $O/e_tests.log 2>&1
echo "tests rc=$?" >> $O/e_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --cpu-seconds 0 > $O/e_bench_weak2.json 2> $O/e_bench_weak2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --cpu-seconds 0 --scaling strong > $O/e_bench_strong2.json 2> $O/e_bench_strong2.err
timeout 300 python bench.py --cpu-seconds 0 --no-predict > $O/e_bench1.json 2> $O/e_bench1.err
tail -n 4 $O/e_tests.log
head -c 400 $O/e_bench_strong2.json
