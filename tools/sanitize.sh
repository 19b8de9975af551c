#!/bin/bash
# compute-sanitizer over the kernels with hand-written synchronisation: the split DMMA path (TMA + mbarrier pipelines,
# warp-specialised transform), the fused sweeps, the small-d kernels, the N-vector reductions and the persistent search
# kernel (grid-wide combinations through the mailbox).  One log per tool under gpurun_out/ (summaries are copied to
# profiles/r2_sanitizer_*.txt).  Usage (GPU box): bash tools/sanitize.sh [memcheck racecheck synccheck]
TOOLS="${@:-memcheck racecheck synccheck}"
SEL='tests/test_gpu_split.py tests/test_gpu_search.py tests/test_gpu_parity.py::test_teacher_forced_steps tests/test_gpu_parity.py::test_reductions_vs_oracle tests/test_gpu_parity.py::test_philox_normal_matches_restatement_and_is_rank_invariant'
mkdir -p gpurun_out
for tool in $TOOLS; do
    log=gpurun_out/r2_sanitizer_$tool.txt
    echo "== compute-sanitizer --tool $tool -- pytest $SEL" > $log
    REE_SANITIZE=1 timeout 1500 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 0 \
        python -m pytest $SEL -m gpu -x -q -p no:cacheprovider >> $log 2>&1
    echo "== exit code $?" >> $log
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|error" $log | tail -5
done
