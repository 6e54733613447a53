#!/usr/bin/env bash
# Builds the C++ slab host driver (plain C ABI, std::thread per rank) -> tests/cpp/build/slab_host.  CPU-only build step.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
mkdir -p "$HERE/build"
/usr/bin/g++ -std=c++17 -O2 -pthread -I"$ROOT/include" "$HERE/slab_host.cpp" \
    -L"$ROOT/fluid_simulator_b200" -lfsim_b200 -Wl,-rpath,"$ROOT/fluid_simulator_b200" -Wl,-rpath,/usr/local/cuda/lib64 -L/usr/local/cuda/lib64 -lcudart \
    -o "$HERE/build/slab_host"
echo "built $HERE/build/slab_host"
