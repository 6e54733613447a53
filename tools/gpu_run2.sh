#!/usr/bin/env bash
# round-2 GPU call B: smem tail without divisions, TMA tile staging of the G2P (A/B), fixed scale-parity tests, launch list
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for cfg in "1 1" "1 0"; do set -- $cfg
  echo "== FSIM_PDL=$1 FSIM_MG_TAIL_SMEM=$2" >> gpurun_out/r2b_projection_ab.log
  FSIM_PDL=$1 FSIM_MG_TAIL_SMEM=$2 timeout 300 python tools/bench_projection.py 64 128 256 >> gpurun_out/r2b_projection_ab.log 2>&1
done
cut -c1-200 gpurun_out/r2b_projection_ab.log
( time timeout 1200 python -m pytest tests -m gpu -q ) > gpurun_out/r2b_tests.log 2>&1
tail -5 gpurun_out/r2b_tests.log
timeout 600 python bench.py --steps 10 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
FSIM_G2P_TMA=0 timeout 600 python bench.py --steps 10 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/r2b_bench_notma.json 2> gpurun_out/r2b_bench_notma.err
FSIM_MG_TAIL_SMEM=0 timeout 600 python bench.py --steps 10 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/r2b_bench_nosmemtail.json 2> gpurun_out/r2b_bench_nosmemtail.err
python - <<'PY'
import json
for f in ("r2b_bench", "r2b_bench_notma", "r2b_bench_nosmemtail"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json")); print(f, d["ms_per_step"], d["config"]["pcg_iterations_mean"], d["kernel_ms"]["advect"], d["checks"]["ok"])
    except Exception as e: print(f, "failed", e)
PY
# every launch of one step with its device time (graph nodes included), to see the tail kernels side by side
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"mg_|spmv|direction|close" -s 400 -c 400 --csv --log-file gpurun_out/r2b_launches_solver.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2b_ncu_launches.log 2>&1
FSIM_MG_TAIL_SMEM=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"mg_tail" -s 40 -c 20 --csv --log-file gpurun_out/r2b_launches_tail_global.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2b_ncu_launches2.log 2>&1
tail -2 gpurun_out/r2b_ncu_launches.log
