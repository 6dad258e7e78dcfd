# full ncu capture of the sensor kernel (E = 1024)
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --envs 1024 --no-cpu-baseline --no-variants"
$CMD > gpurun_out/plain_umma.log 2>&1 && \
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:wfs_frame_umma -c 1 -f -o gpurun_out/prof_wfs_umma_r02_v5 $CMD > gpurun_out/ncu_umma.log 2>&1
tail -2 gpurun_out/ncu_umma.log
