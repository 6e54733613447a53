#!/usr/bin/env bash
# round-2 GPU call N: per-kernel launch list of the cold 256^3 projection (no graph, no PDL) with the current build
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
FSIM_NO_GRAPH=1 FSIM_PDL=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 500 --csv --log-file gpurun_out/r2q_launches_projection.csv \
    python tools/bench_projection.py 256 > gpurun_out/r2q_ncu_launches.log 2>&1
python tools/launch_list_summary.py gpurun_out/r2q_launches_projection.csv | head -30
