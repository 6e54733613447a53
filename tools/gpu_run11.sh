#!/usr/bin/env bash
# round-2 GPU call K: which multigrid levels are visited twice (W legs) -- whole-step effect at 256^3
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {  # name, env...
  local name=$1; shift
  env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/r2l_$name.json 2> gpurun_out/r2l_$name.err
  python - "$name" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/r2l_{n}.json"))
    k = {a: b["ms_per_step"] for a, b in d["kernel_ms"].items()}
    print(n, round(d["ms_per_step"], 3), "ms", d["config"]["pcg_iterations_mean"], "its", d["checks"]["ok"], "mg", k["mg"], "l1", k["mg_level1"], "coarse", k["mg_coarse"])
except Exception as e: print(n, "failed", e)
PY
}
run w2 FSIM_MG_W_FIRST=2 FSIM_MG_W_LAST=2
run none FSIM_MG_W_FIRST=9 FSIM_MG_W_LAST=9
run w3 FSIM_MG_W_FIRST=3 FSIM_MG_W_LAST=3
run w23 FSIM_MG_W_FIRST=2 FSIM_MG_W_LAST=3
run w2b FSIM_MG_W_FIRST=2 FSIM_MG_W_LAST=2
