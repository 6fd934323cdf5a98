set -uo pipefail
mkdir -p gpurun_out
python bench.py > gpurun_out/r01z_bench_B400.log 2>&1; tail -1 gpurun_out/r01z_bench_B400.log | cut -c1-400
bash tools/profile.sh r01z 2>&1 | tail -12
bash tools/traffic.sh r01z 2>&1 | tail -3
python tools/gemm_breakdown.py --out gpurun_out/r01z_gemm_breakdown.json > gpurun_out/r01z_breakdown.log 2>&1; tail -8 gpurun_out/r01z_breakdown.log
