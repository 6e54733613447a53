#!/usr/bin/env bash
# round-2 GPU call D: TMA probe variants, prefetching solver kernels (A/B vs run C's launch list), ncu of the level-0 solver kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/r2d_tma_probe.txt
for v in "32 4 4 4 0" "32 3 4 4 0" "32 -4 4 4 0" "32 -1 4 4 0" "32 4 -1 4 0" "32 4 4 -1 0" "36 4 4 4 0" "40 -4 -1 -1 1" "40 28 3 3 1" "40 60 39 23 1"; do
  timeout 30 tools/_bin/tma_probe $v >> gpurun_out/r2d_tma_probe.txt 2>&1
done
cat gpurun_out/r2d_tma_probe.txt
timeout 300 python tools/bench_projection.py 64 128 256 > gpurun_out/r2d_projection.log 2>&1; cut -c1-200 gpurun_out/r2d_projection.log
FSIM_NO_GRAPH=1 FSIM_PDL=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 500 --csv --log-file gpurun_out/r2d_launches_projection.csv \
    python tools/bench_projection.py 256 > gpurun_out/r2d_ncu_launches.log 2>&1
FSIM_NO_GRAPH=1 FSIM_PDL=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"spmv4|update_first4|jacobi4|restrict4|direction4" -s 40 -c 7 -o gpurun_out/r2d_prof_level0 -f \
    python tools/bench_projection.py 256 > gpurun_out/r2d_ncu_level0.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
python - <<'PY'
import json
for f in ("r2d_bench",):
    try:
        d = json.load(open(f"gpurun_out/{f}.json")); print(f, d["ms_per_step"], d["config"]["pcg_iterations_mean"], d["checks"]["ok"], {k: v["ms_per_step"] for k, v in d["kernel_ms"].items()})
    except Exception as e: print(f, "failed", e)
PY
( time timeout 1200 python -m pytest tests -m gpu -q -x ) > gpurun_out/r2d_tests.log 2>&1
tail -3 gpurun_out/r2d_tests.log
