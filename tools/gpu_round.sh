#!/bin/bash
# One GPU-box visit: parity tests, bench line, iteration/kernel timings, ncu launch list and one full capture.
# Usage (from the repo root, under gpurun): bash tools/gpu_round.sh <tag> [skip-ncu]
set -u
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_$TAG.log
tail -n 5 $OUT/pytest_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -n 2 $OUT/smoke_$TAG.log
timeout 600 python bench.py > $OUT/bench_$TAG.log 2> $OUT/bench_$TAG.err; echo "bench rc=$?"; cat $OUT/bench_$TAG.log
timeout 300 python tools/iter_profile.py --n 10000000 > $OUT/iter_$TAG.log 2>&1; cat $OUT/iter_$TAG.log
timeout 300 python tools/kernel_bench.py --n 4000000 > $OUT/kernels_$TAG.jsonl 2>&1; cat $OUT/kernels_$TAG.jsonl
if [ "${2:-}" != "skip-ncu" ]; then
  CMD="python bench.py --steps 2 --warmup 1 --n 2000000 --no-e2e --no-cpu --no-code-warmup"
  $CMD > $OUT/plain_$TAG.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu_list_$TAG.log 2>&1
  echo "ncu list rc=$?"
  ncu --set full --clock-control none --import-source on -k regex:"k_xform|k_gram2|k_sweep" -s 6 -c 3 -f -o $OUT/prof_$TAG $CMD > $OUT/ncu_full_$TAG.log 2>&1
  echo "ncu full rc=$?"
fi
