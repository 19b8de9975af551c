#!/bin/bash
# Round-2 evidence run (one GPU): full GPU test suite, the bench line, the ncu launch list of the same bench command and
# one `ncu --set full` capture of the three sweep kernels and the search kernel.  Outputs under gpurun_out/.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r2_pytest_gpu.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r2_bench_reference.json 2>> gpurun_out/r2_bench.err
# launch list of the bench command (kernel shares; cold-cache, serialised times)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-code-warmup > gpurun_out/r2_ncu_launches.log 2>&1
# full captures: third launch of each kernel family of the bench command
for k in k_xform_ws k_gram2 "k_xform<" k_adapt; do
  name=$(echo $k | tr -d '<')
  ncu --set full --clock-control none --import-source on -k "regex:$k" -s 2 -c 1 -f -o gpurun_out/r2_$name \
      python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-code-warmup > gpurun_out/r2_ncu_$name.log 2>&1
done
python tools/iter_timeline.py --n 10000000 --iters 2 > gpurun_out/r2_timeline_1e7.txt 2>&1
python tools/iter_timeline.py --n 1250000 --iters 2 > gpurun_out/r2_timeline_1p25e6.txt 2>&1
python tools/config_bench.py pde osc cmc sis > gpurun_out/r2_other_configs_1gpu.jsonl 2>&1
python tools/adapt_bench.py --sizes 1000,1250000,10000000 > gpurun_out/r2_adapt_bench.jsonl 2>&1
tail -3 gpurun_out/r2_pytest_gpu.log; cut -c1-300 gpurun_out/r2_bench.json; cut -c1-400 gpurun_out/r2_bench_reference.json; ls -la gpurun_out/r2_*.ncu-rep
