#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY -- not part of the product.
#
# Compiles the reference Simulator's own translation units, from where they lie under
# $REF_DIR (default /root/reference), into oracle/_ref/ (git-ignored, NOT gpurun-ignored, so the
# built .so files travel to the GPU box, where /root/reference does not exist).
#   oracle/_ref/libfsim_ref.so         -O2 -fopenmp            (timing baseline + parity)
#   oracle/_ref/libfsim_ref_serial.so  -O2 -D_DEBUG  (util/paralellDefine.h:5-10 => every stage
#                                      serial => bit-reproducible single-thread oracle)
# No reference source is copied into the repo: the five TUs are compiled in place; the one-line
# pressure-export patch of bridsonSolverGrid.cpp (the reference keeps `pressure` local,
# bridsonSolverGrid.cpp:249) is applied on the fly through a pipe (sed | g++ -x c++ -).
# The reference's own build system (CMake + Windows-only CPP_LIB_DIR, glfw/spdlog/stb) is not run.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF_DIR="${REF_DIR:-/root/reference}"
SRC="$REF_DIR/src/Simulator/simulator"
OUT="$HERE/_ref"
if [ ! -d "$SRC" ]; then
    echo "build_ref: $SRC not found; keeping prebuilt files in $OUT (if any)" >&2
    exit 0
fi
mkdir -p "$OUT/obj"
CXX="${REF_CXX:-/usr/bin/g++}"  # not $CXX: this image exports CXX=/opt/gcc/bin/g++, a wrapper without libgomp.spec
COMMON="-std=c++20 -O2 -fPIC -w -I$HERE/ref_shim -I$SRC -I$SRC/macGrid -I$SRC/particles"

build_variant() {  # $1 = suffix, $2.. = extra flags
    local sfx="$1"; shift
    local objs=()
    for tu in simulator.cpp macGrid/macGrid.cpp macGrid/basicMacGrid.cpp particles/hashedParticles.cpp; do
        local o="$OUT/obj/$(basename "$tu" .cpp)$sfx.o"
        $CXX $COMMON "$@" -c "$SRC/$tu" -o "$o" &
        objs+=("$o")
    done
    # pressure export: copy the local `pressure` vector out right before it is applied
    local o="$OUT/obj/bridsonSolverGrid$sfx.o"
    sed 's/applyPressureToVelocities(parallel, dt, pressure);/g_ref_last_pressure = pressure; &/' \
        "$SRC/macGrid/bridsonSolverGrid.cpp" |
        $CXX $COMMON "$@" -include "$HERE/ref_patch_decl.h" -x c++ - -c -o "$o" &
    objs+=("$o")
    o="$OUT/obj/ref_harness$sfx.o"
    $CXX $COMMON "$@" -c "$HERE/ref_harness.cpp" -o "$o" &
    objs+=("$o")
    local fail=0
    for j in $(jobs -p); do wait "$j" || fail=1; done
    [ "$fail" = 0 ] || { echo "build_ref: compile failed" >&2; exit 1; }
    $CXX -shared -fopenmp -Wl,-Bsymbolic -o "$OUT/libfsim_ref$sfx.so" "${objs[@]}"
}

build_variant "" -fopenmp
build_variant "_serial" -fopenmp -D_DEBUG
rm -rf "$OUT/obj"
echo "build_ref: built $OUT/libfsim_ref.so and $OUT/libfsim_ref_serial.so"
