for v in 0 12 13 14; do REE_XF_VARIANT=$v python tools/debug/xf_time.py; done
REE_XF_VARIANT=12 timeout 300 python -m pytest tests/test_gpu_split.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -3
