#!/usr/bin/env bash
# Build the UNMODIFIED reference CPU path (plus the two documented compile/UB patches) from
# the sources where they lie under /root/reference into oracle/_ref/.  Test infrastructure
# only: nothing in the product path may execute these binaries.
#
# Patches, applied with sed to scratch copies under a mktemp dir that is
# deleted afterwards (no reference source is ever written into this repo):
#   1. moqui/kernel_functions/mqi_print_data.hpp:22  prints non-existent members
#      (compile error in the non-CUDA branch)       -> print a constant string.
#   2. moqui/base/mqi_p_ionization.hpp:318-320       uint16 wrap-around in the range-table
#      walk (out-of-bounds read -> SEGV)            -> stop the walk at n == 0.
set -euo pipefail
REF=${MQI_REFERENCE:-/root/reference}
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
if [ ! -d "$REF/moqui" ]; then
    echo "build_ref: $REF not present; keeping prebuilt $OUT" >&2
    exit 0
fi
mkdir -p "$OUT"
TMP="$(mktemp -d)"
trap 'rm -rf "$TMP"' EXIT
mkdir -p "$TMP/moqui/kernel_functions" "$TMP/moqui/base"
sed '22s/.*/        printf("%d: material\\n", i);/' \
    "$REF/moqui/kernel_functions/mqi_print_data.hpp" > "$TMP/moqui/kernel_functions/mqi_print_data.hpp"
sed '319s/if (r >= r_steps\[n\]) break;/if (r >= r_steps[n] || n == 0) break;/' \
    "$REF/moqui/base/mqi_p_ionization.hpp" > "$TMP/moqui/base/mqi_p_ionization.hpp"
grep -q 'n == 0' "$TMP/moqui/base/mqi_p_ionization.hpp" || { echo "patch 2 did not apply" >&2; exit 1; }
CXX=${CXX:-g++}
FLAGS="-std=c++11 -O2 -w -DNDEBUG -I$TMP -I$REF"
# phantom_env exactly as the reference's tests/mc/phantom CMake builds it (debug physics) ...
$CXX $FLAGS -D__PHYSICS_DEBUG__ "$REF/tests/mc/phantom/phantom_env.cpp" -o "$OUT/phantom_env_cpu_debug" -lz &
# ... and with the tps CMake's physics (no __PHYSICS_DEBUG__).
$CXX $FLAGS "$REF/tests/mc/phantom/phantom_env.cpp" -o "$OUT/phantom_env_cpu_release" -lz &
# harnesses that drive reference headers directly (KATs, extra scorers, Dij)
for v in debug release; do
    D=""; [ $v = debug ] && D="-D__PHYSICS_DEBUG__"
    $CXX $FLAGS $D "$HERE/ref_kat.cpp" -o "$OUT/ref_kat_$v" -lz &
    $CXX $FLAGS $D "$HERE/ref_harness.cpp" -o "$OUT/ref_harness_$v" -lz &
done
wait
ls -la "$OUT"
