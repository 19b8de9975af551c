for v in 0 3 4 5 6; do REE_XF_VARIANT=$v python bench.py --no-e2e --no-cpu --steps 5 --warmup 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('variant $v', round(d['ms_per_step'],2), {k:round(v['ms_per_launch'],3) for k,v in d['roofline']['kernels'].items()})"; done
