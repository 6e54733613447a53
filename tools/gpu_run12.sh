#!/usr/bin/env bash
# round-2 GPU call L: slab tests on one device after the exchange-kernel changes (PDL launches, wait statistics, owned-plane rhs)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/test_gpu_slab.py tests/test_abi.py -q -x ) > gpurun_out/r2n_tests.log 2>&1
tail -5 gpurun_out/r2n_tests.log
