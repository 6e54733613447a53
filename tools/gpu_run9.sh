#!/usr/bin/env bash
# round-2 GPU call I: out-of-line obstacle paths in the advect kernels; strided export; eager-loading guard
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q -x ) > gpurun_out/r2j_tests.log 2>&1
grep -E "passed|failed|Error" gpurun_out/r2j_tests.log | tail -3
timeout 900 python bench.py --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err
tail -c 300 gpurun_out/r2j_bench.err
timeout 900 python bench.py --steps 10 --warmup 5 --no-cpu-baseline --no-e2e --transfer APIC --obstacle-box > gpurun_out/r2j_bench_apic.json 2> gpurun_out/r2j_bench_apic.err
python - <<'PY'
import json
for f in ("r2j_bench", "r2j_bench_apic"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, round(d["ms_per_step"], 3), d["config"]["pcg_iterations_mean"], d["checks"]["ok"], "e2e", (d.get("e2e") or {}).get("value"), ((d.get("e2e") or {}).get("subsampled_export") or {}).get("value"))
        print("  ", {k: v["ms_per_step"] for k, v in d["kernel_ms"].items()})
    except Exception as e: print(f, "failed", e)
PY
