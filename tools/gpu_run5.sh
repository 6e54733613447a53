#!/usr/bin/env bash
# round-2 GPU call E: aligned TMA box, 2-D blocks for the reducing solver kernels, odd persistent stride, vectorised scan
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/r2e_tma_probe.txt
for v in "40 -4 -1 -1 1" "40 28 3 3 1" "40 60 39 23 1" "40 -4 -1 -1 0"; do timeout 30 tools/_bin/tma_probe $v >> gpurun_out/r2e_tma_probe.txt 2>&1; done
cat gpurun_out/r2e_tma_probe.txt
timeout 300 python tools/bench_projection.py 64 128 256 > gpurun_out/r2e_projection.log 2>&1; cut -c1-200 gpurun_out/r2e_projection.log
( time timeout 1200 python -m pytest tests -m gpu -q -x ) > gpurun_out/r2e_tests.log 2>&1
tail -3 gpurun_out/r2e_tests.log
timeout 600 python bench.py --steps 10 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err
FSIM_G2P_TMA=0 timeout 600 python bench.py --steps 10 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/r2e_bench_notma.json 2> gpurun_out/r2e_bench_notma.err
python - <<'PY'
import json
for f in ("r2e_bench", "r2e_bench_notma"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json")); print(f, d["ms_per_step"], d["config"]["pcg_iterations_mean"], d["checks"]["ok"], {k: v["ms_per_step"] for k, v in d["kernel_ms"].items()})
    except Exception as e: print(f, "failed", e)
PY
FSIM_NO_GRAPH=1 FSIM_PDL=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 500 --csv --log-file gpurun_out/r2e_launches_projection.csv \
    python tools/bench_projection.py 256 > gpurun_out/r2e_ncu_launches.log 2>&1
