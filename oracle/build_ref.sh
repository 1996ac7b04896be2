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
#   3. moqui/base/mqi_file_handler.hpp (scratch copy, used by ref_harness only): the include of
#      mqi_beam_module_ion.hpp (which pulls in the GDCM headers, absent here) is dropped and the file
#      is cut before class file_parser, so that mask_reader (mask_to_roi) compiles on its own.
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
{ sed -n '1,/^class file_parser/p' "$REF/moqui/base/mqi_file_handler.hpp" | sed '$d' \
    | sed '/mqi_beam_module_ion.hpp/d; s/#include <moqui\/base\/mqi_roi.hpp>/#include <moqui\/base\/mqi_roi.hpp>\n#include <moqui\/base\/mqi_common.hpp>\n#include <moqui\/base\/mqi_vec.hpp>\n#include <cassert>\n#include <cstring>\n#include <string>\n#include <vector>/';
  printf '}\n#endif\n'; } > "$TMP/moqui/base/mqi_file_handler.hpp"
grep -q 'mask_to_roi' "$TMP/moqui/base/mqi_file_handler.hpp" || { echo "patch 3 did not apply" >&2; exit 1; }
#   4. the generic part of class file_parser (constructor, read_input_parameters, get_string ... get_bool; the
#      reference file from "class file_parser" up to the line before string_to_scorer_type, which needs the GDCM-side
#      enums) as a scratch header of its own, used by ref_kat section 11 only.
{ printf '#pragma once\n#include <cstring>\n#include <fstream>\n#include <iostream>\n#include <string>\n#include <vector>\n#include <moqui/base/mqi_common.hpp>\n#include <moqui/base/mqi_utils.hpp>\nnamespace mqi {\n';
  sed -n '/^class file_parser/,/string_to_scorer_type/p' "$REF/moqui/base/mqi_file_handler.hpp" | head -n -3;
  printf '};\n}\n'; } > "$TMP/moqui/base/mqi_file_parser_only.hpp"
grep -q 'get_bool' "$TMP/moqui/base/mqi_file_parser_only.hpp" || { echo "patch 4 did not apply" >&2; exit 1; }
#   5. moqui/base/mqi_dataset.hpp, the GDCM-backed DICOM access class, is shadowed by oracle/ref_dataset_stub.hpp (an
#      in-memory stand-in written for this harness; GDCM is absent), so that treatment_machine_pbs / _ion and
#      beam_module_ion compile unmodified for ref_tps_kat.  Nothing else built here includes that header.
cp "$HERE/ref_dataset_stub.hpp" "$TMP/moqui/base/mqi_dataset.hpp"
CXX=${CXX:-g++}
FLAGS="-std=c++11 -O2 -w -DNDEBUG -pthread -I$TMP -I$REF"
# The reference's own CUDA path for sm_100a (nvcc cross-compiles without a GPU): phantom_env.cpp compiled as CUDA
# exactly as tests/mc/phantom/CMakeLists.txt:13-20 does (-w --use_fast_math, source language CUDA), plus
# -maxrregcount=128: the kernel needs 168 registers, so its default 512-thread block does not launch without the
# cap.  ref_harness.cpp compiled the same way drives the same kernel with CUDA events around it
# (GPU baseline of bench.py, high-statistics goldens generated on the GPU box by oracle/gen_golden_gpu.py).
NVCC=${NVCC:-$(command -v nvcc || echo /usr/local/cuda/bin/nvcc)}
if [ -x "$NVCC" ] && [ -z "${MQI_REF_SKIP_CUDA:-}" ]; then
    NVFLAGS="-x cu -std=c++14 -w --use_fast_math -gencode arch=compute_100a,code=sm_100a -maxrregcount=128 -DNDEBUG -I$TMP -I$REF"
    for v in debug release; do
        D=""; [ $v = debug ] && D="-D__PHYSICS_DEBUG__"
        $NVCC $NVFLAGS $D "$REF/tests/mc/phantom/phantom_env.cpp" -o "$OUT/phantom_env_gpu_$v" -lz &
        $NVCC $NVFLAGS $D "$HERE/ref_harness.cpp" -o "$OUT/ref_harness_gpu_$v" -lz &
    done
fi
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
# drop-in proof from the reference's side (ref_dropin.cpp): the reference's own phantom_env, host compiler only, with
# run() routed through the C ABI of libmqi_b200.so as INTEGRATION.md section 1 shows
LIBDIR="$HERE/../moquimc_b200"
if [ -f "$LIBDIR/libmqi_b200.so" ]; then
    for v in debug release; do
        D=""; [ $v = debug ] && D="-D__PHYSICS_DEBUG__"
        $CXX $FLAGS $D -I"$HERE/../include" "$HERE/ref_dropin.cpp" -o "$OUT/ref_dropin_$v" -L"$LIBDIR" -lmqi_b200 \
            -Wl,-rpath,'$ORIGIN/../../moquimc_b200' -lz &
    done
fi
# -O0: treatment_machine_ion::create_beamsource binds a reference to *nullptr for the last spot (characterize_beamlet_time
# ignores it); optimised builds of that undefined behaviour crash
$CXX -std=c++11 -O0 -w -I$TMP -I$REF "$HERE/ref_tps_kat.cpp" -o "$OUT/ref_tps_kat" &
wait
ls -la "$OUT"
