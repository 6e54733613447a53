#!/usr/bin/env bash
# multi-GPU call (gpurun --gpus N): the N-GPU path validated on N real GPUs, then bench.py --gpus N (with its slab pre-flight)
N=${1:-2}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/r2m_n${N}_smi.txt 2>&1
nvidia-smi topo -m > gpurun_out/r2m_n${N}_topo.txt 2>&1
if [ "$SKIP_TESTS" != "1" ]; then ( time timeout 900 python -m pytest tests/test_gpu_slab.py -q -k "real_gpus or two_processes" -s ) > gpurun_out/r2m_n${N}_slabtest.log 2>&1; fi
grep -E "passed|failed|slab worker|rel L2|SLAB_WORKER" gpurun_out/r2m_n${N}_slabtest.log | tail -8
PORT=29512
[ "$SKIP_PCIE" = "1" ] || timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((PORT+7)) tools/pcie_probe_ranks.py > gpurun_out/r2m_n${N}_pcie.json 2> gpurun_out/r2m_n${N}_pcie.err; cut -c1-600 gpurun_out/r2m_n${N}_pcie.json
PYTHONFAULTHANDLER=1 timeout 900 python -u -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $N --steps 20 --warmup 5 \
    > gpurun_out/r2m_n${N}_bench.json 2> gpurun_out/r2m_n${N}_bench.err
echo "bench rc=$?"
tail -c 1500 gpurun_out/r2m_n${N}_bench.err
if [ ! -s gpurun_out/r2m_n${N}_bench.json ]; then
  PYTHONFAULTHANDLER=1 timeout 900 python -u -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((PORT+1)) bench.py --gpus $N --steps 20 --warmup 5 --no-verify \
      > gpurun_out/r2m_n${N}_bench.json 2> gpurun_out/r2m_n${N}_bench_noverify.err
  echo "bench (no verify) rc=$?"
  tail -c 1500 gpurun_out/r2m_n${N}_bench_noverify.err
fi
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2m_n${N}_bench.json"))
    print("N=${N}", round(d["ms_per_step"], 3), "ms/step", d["config"]["pcg_iterations_mean"], "its; checks", d["checks"]["ok"], d["checks"].get("slab_verify"))
    print("   e2e", d["e2e"]["value"] if d.get("e2e") else None, d["e2e"].get("numa") if d.get("e2e") else None)
    print("   ", {k: v["ms_per_step"] for k, v in d["kernel_ms"].items()})
    print("   exchange_wait", d.get("exchange_wait"))
except Exception as e:
    print("bench failed", e)
PY
