mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --envs 1024 --no-cpu-baseline --wfs-path tcgen05"
$CMD > gpurun_out/plain_tc.log 2>&1 && \
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:wfs_frame_tc -c 1 -o gpurun_out/prof_wfs_tc_r01 $CMD > gpurun_out/ncu_tc.log 2>&1
tail -2 gpurun_out/ncu_tc.log
