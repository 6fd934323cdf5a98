#!/usr/bin/env bash
# DRAM traffic of the dominant kernel at the bench configuration (run under gpurun, one GPU):
# ncu dram__bytes_{read,write}.sum of every k_tile_gemm launch; tools/traffic_parse.py keeps the 47 launches of the
# final profiled evaluation (B = 400 GPs per launch) and writes profiles/<tag>_traffic.json.
set -uo pipefail
TAG=${1:-r02}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-predict --cpu-seconds 0"
$CMD > gpurun_out/${TAG}_traffic_plain.log 2>&1 &&
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_tile_gemm \
    --csv --log-file gpurun_out/${TAG}_traffic.csv $CMD > gpurun_out/${TAG}_traffic_ncu.log 2>&1
echo "traffic capture rc=$?"
python tools/traffic_parse.py gpurun_out/${TAG}_traffic.csv gpurun_out/${TAG}_traffic.json
