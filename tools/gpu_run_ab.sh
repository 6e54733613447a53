#!/usr/bin/env bash
# multi-GPU A/B on ONE box: bench.py --gpus N with the all-rank reductions fused into the reducing kernels on / off (FSIM_SLAB_FUSED_AR), no pre-flight
N=${1:-2}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
PORT=29612
for v in ${AB_ORDER:-1 0 1 0}; do
  PORT=$((PORT+3))
  FSIM_SLAB_FUSED_AR=$v timeout 600 python -u -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $N --steps 20 --warmup 5 --no-verify --no-e2e \
      > gpurun_out/r2ab_n${N}_ar$v.json 2> gpurun_out/r2ab_n${N}_ar$v.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2ab_n${N}_ar$v.json"))
    print("N=${N} fused_ar=$v", round(d["ms_per_step"], 3), "ms/step", d["config"]["pcg_iterations_mean"], "its; checks", d["checks"]["ok"])
    print("    ", {k: v["ms_per_step"] for k, v in d["kernel_ms"].items() if k in ("spmv", "pcg_update", "pcg_direction", "mg", "halo", "allreduce")})
    print("    ", {k: (v["ms_per_step_mean_over_ranks"], v["kernels_per_step"], v["in_kernel_ms_per_step_mean_over_ranks"]) for k, v in d["exchange_wait"].items() if k in ("push", "gpush", "allreduce", "fused")})
except Exception as e:
    print("bench failed", e)
PY
done
