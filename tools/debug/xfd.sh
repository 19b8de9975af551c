for v in 7 8 9 10 11; do REE_XF_VARIANT=$v python tools/debug/xf_time.py; done
