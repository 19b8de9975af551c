for dbg in 1 0; do
REE_XF_DEBUG=$dbg REE_XF_VARIANT=4 N=4000000 ncu --set full --clock-control none --import-source on -k regex:k_xform_ws -s 1 -c 1 -f -o gpurun_out/xfws_v4_d$dbg python tools/debug/xf_time.py > gpurun_out/xfws_ncu_d$dbg.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
