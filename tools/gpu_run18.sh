#!/usr/bin/env bash
# round-2: SURVEY cfg 5 (projection only, cold start, tol 1e-6) on one GPU with the final build
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python tools/bench_projection.py 128 192 256 384 512 > gpurun_out/r2_cfg5_projection_sweep.jsonl 2> gpurun_out/r2_cfg5.err
cut -c1-330 gpurun_out/r2_cfg5_projection_sweep.jsonl; tail -2 gpurun_out/r2_cfg5.err
