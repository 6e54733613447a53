#!/usr/bin/env bash
# round-2 GPU call H: one-kernel grid reset, split binning atomic; full bench line (e2e + cpu_baseline), reference arm on the metric's own scene
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q -x ) > gpurun_out/r2i_tests.log 2>&1
grep -E "passed|failed" gpurun_out/r2i_tests.log | tail -2
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err
tail -c 300 gpurun_out/r2i_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2i_bench.json"))
print("bench", round(d["ms_per_step"], 3), d["config"]["pcg_iterations_mean"], d["checks"]["ok"], "e2e", d["e2e"]["value"], "roofline", d["roofline"]["frac"], "cpu", d.get("cpu_baseline", {}).get("value"))
print({k: v["ms_per_step"] for k, v in d["kernel_ms"].items()})
print("step_hbm_frac", d["step_hbm_frac"], d["step_hbm_frac_impl_b_it"])
PY
( time timeout 1500 python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/r2i_bench_reference.json 2> gpurun_out/r2i_bench_reference.err
cut -c1-700 gpurun_out/r2i_bench_reference.json; tail -4 gpurun_out/r2i_bench_reference.err
