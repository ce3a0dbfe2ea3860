#!/bin/sh
# Sanitizer / race-detection targets (the reference has none: SURVEY 5.2).
#
#   tools/sanitize.sh core      ASAN+UBSAN build of the C++ control-plane core, runs tests/test_core.py and
#                               tests/test_ps_round.py against it (CPU)
#   tools/sanitize.sh tsan      ThreadSanitizer run of the cross-GPU flag-barrier protocol model, of the
#                               parameter-server reducer (+ ASAN/UBSAN over the latter) and of the membership-epoch
#                               table under concurrent re-binds (CPU)
#   tools/sanitize.sh memcheck  compute-sanitizer --tool memcheck  over the kernel numerics tests (GPU box)
#   tools/sanitize.sh racecheck compute-sanitizer --tool racecheck (shared-memory hazards; GPU box)
#   tools/sanitize.sh synccheck compute-sanitizer --tool synccheck (barrier misuse; GPU box)
set -eu
ROOT=$(CDPATH= cd -- "$(dirname -- "$0")/.." && pwd)
cd "$ROOT"
PY=${HYPHA_PYTHON:-python}
case "${1:-core}" in
  core)
    OUT=build/asan
    mkdir -p "$OUT"
    EXT=$($PY -c 'import sysconfig; print(sysconfig.get_config_var("EXT_SUFFIX"))')
    INC=$($PY -c 'import pybind11, sysconfig; print("-I", pybind11.get_include(), "-I", sysconfig.get_paths()["include"])')
    # pick a compiler that ships the sanitizer runtimes (the image's $CXX toolchain does not)
    SCXX=${HYPHA_SANITIZE_CXX:-/usr/bin/g++}
    [ -x "$SCXX" ] || SCXX=${CXX:-g++}
    # shellcheck disable=SC2086
    $SCXX -O1 -g -std=c++17 -fPIC -fsanitize=address,undefined -fno-omit-frame-pointer $INC \
        -I hypha_b200/csrc -shared hypha_b200/csrc/core/core_bindings.cpp -o "$OUT/_core$EXT" -pthread
    LIBASAN=$($SCXX -print-file-name=libasan.so)
    # run the reference test vectors against the instrumented module
    LIBCXX=$($SCXX -print-file-name=libstdc++.so.6)   # so ASAN can intercept __cxa_throw
    ASAN_OPTIONS=detect_leaks=0 LD_PRELOAD="$LIBASAN $LIBCXX" HYPHA_CORE_OVERRIDE="$ROOT/$OUT" \
        $PY -m pytest tests/test_core.py tests/test_ps_round.py -q -p no:cacheprovider
    ;;
  tsan)
    unset LD_PRELOAD ASAN_OPTIONS || true      # when called from the ASAN run of the test-suite (`core` target)
    OUT=build/tsan
    mkdir -p "$OUT"
    SCXX=${HYPHA_SANITIZE_CXX:-/usr/bin/g++}
    [ -x "$SCXX" ] || SCXX=${CXX:-g++}
    $SCXX -O1 -g -std=c++17 -fsanitize=thread -fno-omit-frame-pointer -I hypha_b200/csrc \
        tools/tsan_flag_barrier.cpp -o "$OUT/tsan_flag_barrier" -pthread
    "$OUT/tsan_flag_barrier"
    $SCXX -O1 -g -std=c++17 -fsanitize=thread -fno-omit-frame-pointer -I hypha_b200/csrc \
        tools/tsan_ps_reduce.cpp -o "$OUT/tsan_ps_reduce" -pthread
    "$OUT/tsan_ps_reduce"
    # membership-epoch table of the fused outer step: snapshots under concurrent re-binds
    $SCXX -O1 -g -std=c++17 -fsanitize=thread -fno-omit-frame-pointer -I hypha_b200/csrc \
        tools/tsan_epoch_table.cpp -o "$OUT/tsan_epoch_table" -pthread
    "$OUT/tsan_epoch_table"
    # the same driver under ASAN+UBSAN: mmap bounds and unaligned tensor data
    $SCXX -O1 -g -std=c++17 -fsanitize=address,undefined -fno-omit-frame-pointer -I hypha_b200/csrc \
        tools/tsan_ps_reduce.cpp -o "$OUT/asan_ps_reduce" -pthread
    ASAN_OPTIONS=detect_leaks=1 "$OUT/asan_ps_reduce"
    ;;
  memcheck|racecheck|synccheck)
    command -v compute-sanitizer >/dev/null 2>&1 || { echo "compute-sanitizer not found" >&2; exit 1; }
    HYPHA_NO_AUTOBUILD=1 compute-sanitizer --tool "$1" --error-exitcode 1 \
        $PY -m pytest tests/test_kernels_gpu.py -m gpu -q -x \
        -k "rmsnorm or rope or swiglu or cross_entropy or embedding or adamw or outer_step_reference or gemm_accumulate or gemm_strided or attention_fwd_tcgen05 or attention_bwd_tcgen05"
    ;;
  *) echo "usage: $0 {core|tsan|memcheck|racecheck|synccheck}" >&2; exit 2 ;;
esac
