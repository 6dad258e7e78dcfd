mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --envs 4096 --no-cpu-baseline"
$CMD > gpurun_out/plain_gtc.log 2>&1 && \
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -c 8 -o gpurun_out/prof_gemmtc_r01 $CMD > gpurun_out/ncu_gtc.log 2>&1
tail -2 gpurun_out/ncu_gtc.log
