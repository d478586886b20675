#!/bin/sh
# Stage the reference package for runs on the GPU box: /root/reference does not exist there, so the
# unmodified `causarray` package is copied into the git-ignored oracle/_ref/ (it travels with the
# gpurun snapshot like the built .so files; it never enters the history).  Used by
# bench.py --impl reference and by the full-size parity tools through oracle/ref_shim.py.
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
SRC="${CAUSARRAY_REFERENCE_ROOT:-/root/reference}"
if [ ! -d "$SRC/causarray" ]; then
  echo "stage_ref: no reference tree at $SRC (nothing staged)"; exit 0
fi
mkdir -p "$ROOT/oracle/_ref"
rm -rf "$ROOT/oracle/_ref/causarray"
cp -r "$SRC/causarray" "$ROOT/oracle/_ref/causarray"
find "$ROOT/oracle/_ref" -name __pycache__ -type d -prune -exec rm -rf {} +
( cd "$SRC" && git rev-parse HEAD 2>/dev/null || echo unknown ) > "$ROOT/oracle/_ref/SOURCE_REV"
echo "stage_ref: staged $SRC/causarray -> oracle/_ref/causarray"
