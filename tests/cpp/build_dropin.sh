#!/usr/bin/env bash
# Drop-in demonstration (only where /root/reference exists): the reference's OWN manager/simulationManager.{h,cpp} -- the
# caller of the hot path -- is compiled UNMODIFIED against the facade headers and linked with libfsim_b200.so, plus a tiny
# headless driver.  Sources are piped (cat | g++ -x c++ -) so that their `../simulator/...` includes resolve to
# fluid_simulator_b200/host/simulator and no reference source is copied.  Output: oracle/_ref/manager_on_b200
# (a binary; oracle/_ref is git-ignored and travels to the GPU box).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
REF="${REF_DIR:-/root/reference}/src/Simulator/manager"
HOST="$ROOT/fluid_simulator_b200/host"
if [ ! -f "$REF/simulationManager.cpp" ]; then echo "build_dropin: reference not present, skipping"; exit 0; fi
mkdir -p "$ROOT/oracle/_ref"
TMP="$(mktemp -d)"
trap 'rm -rf "$TMP"' EXIT
( cat "$REF/simulationManager.h"; grep -v '#include "simulationManager.h"' "$REF/simulationManager.cpp"; cat "$HERE/manager_drive.inc" ) |
    ( cd "$TMP" && /usr/bin/g++ -std=c++20 -O2 -fopenmp -w -I"$HOST/manager" -I"$HOST" -x c++ - -c -o "$TMP/manager.o" )
/usr/bin/g++ -std=c++20 -O2 -fopenmp -I"$HOST" -c "$HOST/facade.cpp" -o "$TMP/facade.o"
/usr/bin/g++ -fopenmp "$TMP/manager.o" "$TMP/facade.o" -L"$ROOT/fluid_simulator_b200" -lfsim_b200 -Wl,-rpath,"$ROOT/fluid_simulator_b200" \
    -Wl,-rpath,/usr/local/cuda/lib64 -L/usr/local/cuda/lib64 -lcudart -lpthread -o "$ROOT/oracle/_ref/manager_on_b200"
echo "built $ROOT/oracle/_ref/manager_on_b200"
