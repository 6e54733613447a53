#!/usr/bin/env bash
# round-2 GPU call C: TMA probe, ncu of the two multigrid tail kernels, launch list of one cold projection, GPU test-suite
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 60 tools/_bin/tma_probe > gpurun_out/r2c_tma_probe.txt 2>&1; cat gpurun_out/r2c_tma_probe.txt
FSIM_NO_GRAPH=1 FSIM_PDL=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mg_tail" -s 6 -c 2 -o gpurun_out/r2c_prof_tail_smem -f \
    python tools/bench_projection.py 256 > gpurun_out/r2c_ncu_tail_smem.log 2>&1
FSIM_NO_GRAPH=1 FSIM_PDL=0 FSIM_MG_TAIL_SMEM=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mg_tail" -s 6 -c 2 -o gpurun_out/r2c_prof_tail_global -f \
    python tools/bench_projection.py 256 > gpurun_out/r2c_ncu_tail_global.log 2>&1
FSIM_NO_GRAPH=1 FSIM_PDL=0 FSIM_MG_TAIL_SMEM=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 500 --csv --log-file gpurun_out/r2c_launches_projection.csv \
    python tools/bench_projection.py 256 > gpurun_out/r2c_ncu_launches.log 2>&1
( time timeout 1200 python -m pytest tests -m gpu -q ) > gpurun_out/r2c_tests.log 2>&1
tail -5 gpurun_out/r2c_tests.log
