#!/usr/bin/env bash
# round-2 GPU call A: full GPU test-suite, bench line of the new bench.py, ncu captures of the particle kernels, 256^3 parity
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt 2>&1
nproc > gpurun_out/r2a_host.txt; free -g >> gpurun_out/r2a_host.txt
timeout 120 tools/_bin/red_probe > gpurun_out/r2a_red_probe.txt 2>&1; cat gpurun_out/r2a_red_probe.txt
# solver A/B first (seconds each): programmatic dependent launch and the shared-memory-resident multigrid tail, on and off
for cfg in "1 1" "0 1" "1 0" "0 0"; do set -- $cfg
  echo "== FSIM_PDL=$1 FSIM_MG_TAIL_SMEM=$2" >> gpurun_out/r2a_projection_ab.log
  FSIM_PDL=$1 FSIM_MG_TAIL_SMEM=$2 timeout 300 python tools/bench_projection.py 64 128 256 >> gpurun_out/r2a_projection_ab.log 2>&1
done
cat gpurun_out/r2a_projection_ab.log | cut -c1-220
( time timeout 1200 python -m pytest tests -m gpu -q ) > gpurun_out/r2a_tests.log 2>&1
tail -5 gpurun_out/r2a_tests.log
timeout 600 python bench.py --steps 10 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
tail -c 600 gpurun_out/r2a_bench.err
FSIM_PDL=0 timeout 600 python bench.py --steps 10 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/r2a_bench_nopdl.json 2> gpurun_out/r2a_bench_nopdl.err
tail -c 300 gpurun_out/r2a_bench_nopdl.err
FSIM_MG_TAIL_SMEM=0 timeout 600 python bench.py --steps 10 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/r2a_bench_nosmemtail.json 2> gpurun_out/r2a_bench_nosmemtail.err
FSIM_PDL=0 FSIM_MG_TAIL_SMEM=0 timeout 600 python bench.py --steps 10 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/r2a_bench_r1solver.json 2> gpurun_out/r2a_bench_r1solver.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"p2g_kernel|g2p_advect_kernel|reorder_kernel" -s 5 -c 3 \
    -o gpurun_out/r2a_prof_particles -f python bench.py --steps 1 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/r2a_ncu.log 2>&1
tail -3 gpurun_out/r2a_ncu.log
( time timeout 1500 python tools/parity_256.py --tol 1e-6 --out gpurun_out/r2a_parity_256.json ) > gpurun_out/r2a_parity_256.log 2>&1
tail -4 gpurun_out/r2a_parity_256.log
