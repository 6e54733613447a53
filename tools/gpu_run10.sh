#!/usr/bin/env bash
# round-2 GPU call J: fused CG kernels in the hybrid slab projection (one GPU, in-process slab groups) + driver-argument bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/r2k_tests.log 2>&1
grep -E "passed|failed|Error" gpurun_out/r2k_tests.log | tail -3
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err
tail -c 300 gpurun_out/r2k_bench.err
python - <<'PY'
import json
for f in ("r2k_bench",):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, round(d["ms_per_step"], 3), d["config"]["pcg_iterations_mean"], d["checks"]["ok"], "e2e", (d.get("e2e") or {}).get("value"), ((d.get("e2e") or {}).get("subsampled_export") or {}).get("value"))
        print("  ", {k: v["ms_per_step"] for k, v in d["kernel_ms"].items()})
        print("  roofline", d["roofline"], "cpu", d.get("cpu_baseline"))
    except Exception as e: print(f, "failed", e)
PY
