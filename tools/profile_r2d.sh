#!/bin/bash
# round-2 capture of the small-d sweeps (oscillator share of BASELINE configs[3]: d = 6, N = 1.25e7 on one GPU):
# one `ncu --set full` launch of k_small_step_g / k_small_post_g each, summarised on the box (the .ncu-rep stay in /tmp).
mkdir -p gpurun_out
B="python bench.py --n 12500000 --d 6 --steps 2 --warmup 3 --no-e2e --no-cpu --no-code-warmup --no-fp64-probe"
$B > gpurun_out/r2_small_bench.json 2> gpurun_out/r2_small_bench.err || exit 1
: > gpurun_out/r2_small_d_kernels_ncu.txt
for k in k_small_step_g k_small_post_g; do
  ncu --set full --clock-control none --import-source on -k "regex:^$k" -s 2 -c 1 -f -o /tmp/r2_$k $B > /tmp/ncu_$k.log 2>&1
  ncu -i /tmp/r2_$k.ncu-rep --page raw --csv > /tmp/r2_$k.csv 2>/dev/null
  echo "== $k  (ncu --set full --clock-control none, third launch of: $B)" >> gpurun_out/r2_small_d_kernels_ncu.txt
  python tools/ncu_raw_summary.py /tmp/r2_$k.csv >> gpurun_out/r2_small_d_kernels_ncu.txt
done
cut -c1-300 gpurun_out/r2_small_bench.json; cat gpurun_out/r2_small_d_kernels_ncu.txt
