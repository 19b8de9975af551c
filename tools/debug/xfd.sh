for v in 23 24 25 26; do REE_XF_VARIANT=$v timeout 120 python tools/debug/xf_time.py; done
