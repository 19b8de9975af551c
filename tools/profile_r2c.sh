#!/bin/bash
# final evidence refresh of round 2 (one GPU): bench lines, launch list, full captures of the changed sweep kernels
# (exported to CSV on the box: gpurun brings back at most 64 MiB)
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench_reference.json 2>> gpurun_out/r2_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-code-warmup --no-fp64-probe > gpurun_out/r2_ncu_launches.log 2>&1
for k in k_xform_ws '^k_xform$'; do
  name=$(echo $k | tr -d '^$')
  ncu --set full --clock-control none --import-source on -k "regex:$k" -s 2 -c 1 -f -o /tmp/r2_$name \
      python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-code-warmup --no-fp64-probe > gpurun_out/r2_ncu_$name.log 2>&1
  ncu -i /tmp/r2_$name.ncu-rep --page raw --csv > gpurun_out/r2_raw_$name.csv 2>/dev/null
done
python tools/iter_timeline.py --n 10000000 --iters 2 > gpurun_out/r2_timeline_1e7.txt 2>&1
python tools/iter_timeline.py --n 1250000 --iters 2 > gpurun_out/r2_timeline_1p25e6.txt 2>&1
python tools/config_bench.py pde osc cmc sis > gpurun_out/r2_other_configs_1gpu.jsonl 2>&1
cut -c1-260 gpurun_out/r2_bench.json; du -sh gpurun_out
