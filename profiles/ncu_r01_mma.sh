mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --envs 1024 --no-cpu-baseline"
$CMD > gpurun_out/plain_mma.log 2>&1 && \
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:wfs_frame_mma -c 1 -o gpurun_out/prof_wfs_mma_r01 $CMD > gpurun_out/ncu_mma.log 2>&1
ls -la gpurun_out; tail -c 600 gpurun_out/plain_mma.log
