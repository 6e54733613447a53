#!/usr/bin/env bash
# round-2 GPU call F: tail kernel with block-local coarse levels (A/B), 4-wide extrapolation, dynamic smem G2P staging (TMA A/B)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in 1 0; do echo "== FSIM_MG_TAIL2=$v" >> gpurun_out/r2f_projection.log; FSIM_MG_TAIL2=$v timeout 300 python tools/bench_projection.py 64 128 256 >> gpurun_out/r2f_projection.log 2>&1; done
cut -c1-200 gpurun_out/r2f_projection.log
( time timeout 1200 python -m pytest tests -m gpu -q -x ) > gpurun_out/r2f_tests.log 2>&1
grep -E "passed|failed" gpurun_out/r2f_tests.log | tail -2
timeout 600 python bench.py --steps 10 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
FSIM_G2P_TMA=0 timeout 600 python bench.py --steps 10 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/r2f_bench_notma.json 2> gpurun_out/r2f_bench_notma.err
FSIM_MG_TAIL2=0 timeout 600 python bench.py --steps 10 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/r2f_bench_notail2.json 2> gpurun_out/r2f_bench_notail2.err
timeout 600 python bench.py --steps 10 --warmup 5 --no-cpu-baseline --no-e2e --transfer APIC --obstacle-box > gpurun_out/r2f_bench_apic.json 2> gpurun_out/r2f_bench_apic.err
FSIM_G2P_TMA=0 timeout 600 python bench.py --steps 10 --warmup 5 --no-cpu-baseline --no-e2e --transfer APIC --obstacle-box > gpurun_out/r2f_bench_apic_notma.json 2> gpurun_out/r2f_bench_apic_notma.err
python - <<'PY'
import json
for f in ("r2f_bench", "r2f_bench_notma", "r2f_bench_notail2", "r2f_bench_apic", "r2f_bench_apic_notma"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json")); print(f, round(d["ms_per_step"], 3), d["config"]["pcg_iterations_mean"], d["checks"]["ok"], {k: v["ms_per_step"] for k, v in d["kernel_ms"].items()})
    except Exception as e: print(f, "failed", e)
PY
