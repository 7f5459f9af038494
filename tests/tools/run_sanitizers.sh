#!/bin/sh
# Memory-safety check without a GPU (compute-sanitizer is not available on the GPU pool): the kernel
# source compiled for the host (tests/emu/) and the host text pipeline (csrc/grimb_text.cpp) are
# rebuilt with -fsanitize=address,undefined and driven by the golden cases, the random-table cases,
# the host-logic tests and the garbage-input fuzz.  Any sanitizer report fails the run.
set -e
ROOT="$(cd "$(dirname "$0")/../.." && pwd)"
OUT="${TMPDIR:-/tmp}/grimb_sanitize"
mkdir -p "$OUT"
SAN="-O1 -g -std=c++17 -ffp-contract=off -fPIC -shared -pthread -fsanitize=address,undefined -fno-omit-frame-pointer"
g++ $SAN -DGRIMB_EMU -Wno-unused-variable -o "$OUT/libgrimb_emu_asan.so" "$ROOT/tests/emu/grimb_emu.cpp"
g++ $SAN -I"$ROOT/include" -DGRIMB_KW=1 -DGRIMB_KEY_WORDS=1 -o "$OUT/libgrimb_text_asan.so" \
    "$ROOT/py-graph-imputation_b200/csrc/grimb_text.cpp" "$ROOT/tests/tools/text_stubs.cpp"
export GRIMB_EMU_SO="$OUT/libgrimb_emu_asan.so" GRIMB_LIB="$OUT/libgrimb_text_asan.so"
export LD_PRELOAD="$(gcc -print-file-name=libasan.so)" ASAN_OPTIONS=detect_leaks=0:halt_on_error=1
cd "$ROOT"
python -m pytest tests/test_emu_golden.py tests/test_emu_random_tables.py tests/test_text_pipeline.py \
    tests/test_host_logic.py -x -q -s > "$OUT/pytest.log" 2>&1 || { tail -30 "$OUT/pytest.log"; exit 1; }
python tests/golden/fuzz_garbage_text.py 3 60 > "$OUT/garbage.log" 2>&1 || { tail -30 "$OUT/garbage.log"; exit 1; }
if grep -E "runtime error|AddressSanitizer" "$OUT/pytest.log" "$OUT/garbage.log"; then
  echo "SANITIZER REPORTS FOUND"; exit 1
fi
tail -1 "$OUT/pytest.log"; tail -1 "$OUT/garbage.log"; echo "sanitizers: clean"
