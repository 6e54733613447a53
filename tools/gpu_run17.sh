#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2h_bench.json"))
print(d["ms_per_step"], d["e2e"]["value"], d["e2e"]["subsampled_export"])
PY
