#!/usr/bin/env bash
# round-2 GPU call M: four-cells-per-thread kernels for the stored-operator multigrid levels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/r2p_tests.log 2>&1
grep -E "passed|failed|Error" gpurun_out/r2p_tests.log | tail -3
grep coarse_c4 gpurun_out/parity_diag.jsonl | tail -2
run() {
  local name=$1; shift
  env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/r2p_$name.json 2> gpurun_out/r2p_$name.err
  python - "$name" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/r2p_{n}.json"))
    k = {a: b["ms_per_step"] for a, b in d["kernel_ms"].items()}
    print(n, round(d["ms_per_step"], 3), "ms", d["config"]["pcg_iterations_mean"], "its", d["checks"]["ok"], "mg", k["mg"], "l1", k["mg_level1"], "coarse", k["mg_coarse"])
except Exception as e: print(n, "failed", e)
PY
}
run c4_default FSIM_MG_C4_MIN=200000
run c4_off FSIM_MG_C4_MIN=999999999999
run c4_l1only FSIM_MG_C4_MIN=1000000
run c4_default2 FSIM_MG_C4_MIN=200000
timeout 300 python tools/bench_projection.py 128 256 > gpurun_out/r2p_projection.log 2>&1; cut -c1-300 gpurun_out/r2p_projection.log | tail -4
