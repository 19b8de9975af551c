#!/bin/bash
# multi-GPU evidence (run with gpurun --gpus G): strong-scaling bench line, the 8-GPU configurations at full size,
# multi-GPU parity tests.  usage: bash tools/scale_r2.sh G [configs] [tests]
G=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1"
mkdir -p gpurun_out
$TR --master-port 29601 bench.py --gpus $G --steps 10 --warmup 3 > gpurun_out/r2_scale_$G.json 2> gpurun_out/r2_scale_$G.err
tail -1 gpurun_out/r2_scale_$G.json | cut -c1-400
if [ "$2" = "configs" ]; then
  $TR --master-port 29602 tools/config_bench.py osc cmc sis > gpurun_out/r2_configs_${G}gpu.txt 2>&1
  grep '^{' gpurun_out/r2_configs_${G}gpu.txt | cut -c1-330
fi
if [ "$3" = "tests" ]; then
  timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -4
fi
