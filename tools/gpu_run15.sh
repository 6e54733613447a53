#!/usr/bin/env bash
# round-2 final single-GPU evidence: full GPU tests, the driver-argument bench line, the reference arm on 256^3 is NOT repeated here
# (profiles/r2_bench_reference_256.json), launch list of one step (no graph, no PDL: ncu cannot see the nodes of a conditional
# graph), --set full captures of the particle kernels and of the level-0 / level-1 solver kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r2f_tests.log 2>&1
grep -E "passed|failed|Error" gpurun_out/r2f_tests.log | tail -3
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
tail -c 300 gpurun_out/r2f_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2f_bench.json"))
print("bench", round(d["ms_per_step"], 3), d["value"], d["config"]["pcg_iterations_mean"], d["checks"]["ok"], "e2e", d["e2e"]["value"], d["e2e"]["subsampled_export"]["value"])
print("  ", {k: v["ms_per_step"] for k, v in d["kernel_ms"].items()})
print("  roofline", d["roofline"]["frac"], d["roofline"]["avg_launch_ms"], "step frac", d["step_hbm_frac"], "cpu", d.get("cpu_baseline"))
PY
FSIM_NO_GRAPH=1 FSIM_PDL=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2f_launches_256.csv \
    python bench.py --steps 1 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/r2f_ncu_launches.log 2>&1
python tools/launch_list_summary.py gpurun_out/r2f_launches_256.csv | head -45
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"p2g_kernel|g2p_advect_kernel|reorder_kernel" -s 5 -c 3 \
    -o gpurun_out/r2f_prof_particles -f python bench.py --steps 1 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/r2f_ncu.log 2>&1
tail -2 gpurun_out/r2f_ncu.log
python tools/ncu_summary.py gpurun_out/r2f_prof_particles.ncu-rep > gpurun_out/r2f_ncu_particles_summary.txt 2>&1; head -30 gpurun_out/r2f_ncu_particles_summary.txt
FSIM_NO_GRAPH=1 FSIM_PDL=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"c4_kernel|mg_jacobi4_kernel|mg_update_first4|spmv4" -s 60 -c 8 \
    -o gpurun_out/r2f_prof_solver -f python tools/bench_projection.py 256 > gpurun_out/r2f_ncu2.log 2>&1
python tools/ncu_summary.py gpurun_out/r2f_prof_solver.ncu-rep > gpurun_out/r2f_ncu_solver_summary.txt 2>&1; head -40 gpurun_out/r2f_ncu_solver_summary.txt
ls -la gpurun_out/*.ncu-rep
