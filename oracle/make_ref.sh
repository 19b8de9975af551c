#!/bin/sh
# TEST / BENCH INFRASTRUCTURE ONLY.  Recipe for oracle/_ref: an unmodified copy of the reference's Python package
# (pure Python -- nothing to compile) taken from the sources where they lie under /root/reference.  oracle/_ref/ is
# git-ignored (it never enters the history) but not gpurun-ignored, so it travels to the GPU box like the built .so:
# there `bench.py --impl reference` times the reference's own CBREE.solve (solver.py:365-503) on the host cores and the
# `needs_reference` tests can run.  Usage: sh oracle/make_ref.sh [reference-root]   (default /root/reference)
set -e
REF="${1:-/root/reference}"
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
if [ ! -d "$REF/src/rareeventestimation" ]; then
    echo "make_ref.sh: $REF/src/rareeventestimation not found; leaving $OUT as it is" >&2
    exit 0
fi
rm -rf "$OUT"
mkdir -p "$OUT/src"
# the whole package tree (94 files, 6.6 kLoC of .py): solver / utilities / solution / mixturemodel / problem / sis / era /
# evaluation are all reached from CBREE.solve, SIS, CMC or do_multiple_solves
cp -r "$REF/src/rareeventestimation" "$OUT/src/rareeventestimation"
find "$OUT" -name '__pycache__' -type d -prune -exec rm -rf {} +
find "$OUT" -name '*.ipynb' -delete
( cd "$REF" && git rev-parse HEAD 2>/dev/null || echo "unknown" ) > "$OUT/REVISION"
echo "oracle/_ref: $(find "$OUT" -name '*.py' | wc -l) python files copied from $REF"
