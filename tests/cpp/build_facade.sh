#!/usr/bin/env bash
# Builds the facade smoke driver (C++ host side over the C ABI) -> tests/cpp/build/facade_smoke.  CPU-only build step.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
HOST="$ROOT/fluid_simulator_b200/host"
mkdir -p "$HERE/build"
/usr/bin/g++ -std=c++20 -O2 -fopenmp -I"$HOST" "$HERE/facade_smoke.cpp" "$HOST/facade.cpp" \
    -L"$ROOT/fluid_simulator_b200" -lfsim_b200 -Wl,-rpath,"$ROOT/fluid_simulator_b200" -Wl,-rpath,/usr/local/cuda/lib64 -L/usr/local/cuda/lib64 -lcudart \
    -o "$HERE/build/facade_smoke"
echo "built $HERE/build/facade_smoke"
# reconfiguration paths of the drop-in surface (setParticleNum / setParticleR / updateGridParams + setNewMacGrid / restart)
/usr/bin/g++ -std=c++20 -O2 -fopenmp -I"$HOST" "$HERE/facade_reconfig.cpp" "$HOST/facade.cpp" \
    -L"$ROOT/fluid_simulator_b200" -lfsim_b200 -Wl,-rpath,"$ROOT/fluid_simulator_b200" -Wl,-rpath,/usr/local/cuda/lib64 -L/usr/local/cuda/lib64 -lcudart \
    -o "$HERE/build/facade_reconfig"
echo "built $HERE/build/facade_reconfig"
