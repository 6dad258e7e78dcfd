mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --envs 1024 --no-cpu-baseline --no-variants --wfs-path umma_ws"
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:wfs_frame_ws -c 1 -f -o gpurun_out/prof_wfs_ws_r02_v1 $CMD > gpurun_out/ncu_ws.log 2>&1
tail -2 gpurun_out/ncu_ws.log
