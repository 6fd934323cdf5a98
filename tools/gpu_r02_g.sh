#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
GPRB200_LIB=$PWD/gpr.jl_b200/libgprb200_oldgemm.so timeout 300 python bench.py --cpu-seconds 0 --no-predict > $O/g_bench_oldgemm.json 2> $O/g_bench_oldgemm.err
timeout 300 python bench.py --cpu-seconds 0 --no-predict > $O/g_bench_new.json 2> $O/g_bench_new.err
GPRB200_LIB=$PWD/gpr.jl_b200/libgprb200_oldgemm.so timeout 300 python bench.py --trials 13 --steps 10 --cpu-seconds 0 --no-predict > $O/g_bench_oldgemm13.json 2>> $O/g_bench_oldgemm.err
