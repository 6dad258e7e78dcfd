mkdir -p gpurun_out
CMD="python profiles/dev/denoise_tc_probe.py 148000"
AOM_DN_TIMING=1 $CMD > gpurun_out/plain_dntc.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:denoise_tc_kernel -s 14 -c 1 -o gpurun_out/prof_denoise_tc_r02 $CMD > gpurun_out/ncu_dntc.log 2>&1
grep -A8 "load+e1" gpurun_out/plain_dntc.log | tail -9; tail -4 gpurun_out/plain_dntc.log; tail -2 gpurun_out/ncu_dntc.log
