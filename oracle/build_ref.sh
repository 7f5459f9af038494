#!/bin/sh
# TEST INFRASTRUCTURE: makes the UNMODIFIED reference importable from oracle/_ref/ (git-ignored, but it
# travels to the GPU box with the snapshot, like the in-tree .so files), for bench.py --impl reference and
# for validating oracle/grim_oracle.py.  Run where /root/reference exists (the build container); elsewhere the
# prebuilt oracle/_ref is used as is.  Nothing of the reference enters the repository history.
#   - copies the reference's `grim` package (pure Python + cutils.pyx) into oracle/_ref/grim
#   - compiles grim/imputation/cutils.pyx with Cython + gcc (what the reference's setup.py:88-96 does)
set -e
REF="${1:-/root/reference}"
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
[ -d "$REF/grim" ] || { echo "build_ref.sh: $REF/grim not found (keeping any prebuilt $OUT)"; exit 0; }
if [ -f "$OUT/.stamp" ] && [ "$OUT/.stamp" -nt "$REF/grim/imputation/impute.py" ]; then exit 0; fi
rm -rf "$OUT"
mkdir -p "$OUT"
cp -r "$REF/grim" "$OUT/grim"
find "$OUT" -name "__pycache__" -type d -prune -exec rm -rf {} +
cd "$OUT/grim/imputation"
PY="${PYTHON:-python}"
"$PY" -m cython -3 cutils.pyx -o cutils.c
INC="$("$PY" -c 'import sysconfig; print(sysconfig.get_paths()["include"])')"
SUF="$("$PY" -c 'import sysconfig; print(sysconfig.get_config_var("EXT_SUFFIX"))')"
gcc -O2 -fPIC -shared -I"$INC" cutils.c -o "cutils$SUF"
rm -f cutils.c
touch "$OUT/.stamp"
echo "build_ref.sh: reference vendored into $OUT"
