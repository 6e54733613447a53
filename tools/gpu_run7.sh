#!/usr/bin/env bash
# round-2 GPU call G: active-tile list for the level-0 Jacobi sweep (A/B via FSIM_MG_TILES), launch list
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in 1 0; do echo "== FSIM_MG_TILES=$v" >> gpurun_out/r2g_projection.log; FSIM_MG_TILES=$v timeout 300 python tools/bench_projection.py 64 128 256 >> gpurun_out/r2g_projection.log 2>&1; done
cut -c1-200 gpurun_out/r2g_projection.log
FSIM_NO_GRAPH=1 FSIM_PDL=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 500 --csv --log-file gpurun_out/r2g_launches_projection.csv \
    python tools/bench_projection.py 256 > gpurun_out/r2g_ncu_launches.log 2>&1
python tools/launch_list_summary.py gpurun_out/r2g_launches_projection.csv --iteration | tail -26
( time timeout 1200 python -m pytest tests -m gpu -q -x ) > gpurun_out/r2g_tests.log 2>&1
grep -E "passed|failed" gpurun_out/r2g_tests.log | tail -2
timeout 600 python bench.py --steps 10 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
python - <<'PY'
import json
for f in ("r2g_bench",):
    try:
        d = json.load(open(f"gpurun_out/{f}.json")); print(f, round(d["ms_per_step"], 3), d["config"]["pcg_iterations_mean"], d["checks"]["ok"], {k: v["ms_per_step"] for k, v in d["kernel_ms"].items()})
    except Exception as e: print(f, "failed", e)
PY
