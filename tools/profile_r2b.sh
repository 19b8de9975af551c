#!/bin/bash
# second part of the round-2 evidence run: launch list without the fp64 probe, Mahalanobis kernel capture, timelines
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-code-warmup --no-fp64-probe > gpurun_out/r2_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:^k_xform$' -s 2 -c 1 -f -o gpurun_out/r2_k_xform \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-code-warmup --no-fp64-probe > gpurun_out/r2_ncu_k_xform.log 2>&1
python tools/iter_timeline.py --n 10000000 --iters 2 > gpurun_out/r2_timeline_1e7.txt 2>&1
python tools/iter_timeline.py --n 1250000 --iters 2 > gpurun_out/r2_timeline_1p25e6.txt 2>&1
python tools/config_bench.py pde osc cmc sis > gpurun_out/r2_other_configs_1gpu.jsonl 2>&1
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_bench_b.json 2>/dev/null
tail -12 gpurun_out/r2_timeline_1p25e6.txt; cut -c1-200 gpurun_out/r2_bench_b.json; ls gpurun_out/r2_k_xform.ncu-rep
