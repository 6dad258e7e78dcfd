mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --envs 1024 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01.csv $CMD > gpurun_out/ncu_list.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:wfs_frame -c 2 -o gpurun_out/prof_wfs_r01 $CMD > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out; tail -2 gpurun_out/plain.log
