#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python tools/step_trace.py 256 90 > gpurun_out/r2g_step_trace_256.jsonl 2> gpurun_out/r2g_step_trace.err
tail -3 gpurun_out/r2g_step_trace.err
python - <<'PY'
import json
for ln in open("gpurun_out/r2g_step_trace_256.jsonl"):
    d = json.loads(ln)
    if d["step"] % 3 == 0 or d["ms"] > 20: print(d["step"], d["ms"], d["its"], "%.2e" % d["rmax"], d["stages_us"])
PY
