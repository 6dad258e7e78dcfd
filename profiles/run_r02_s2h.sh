mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-variants"
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel" -s 3 -c 3 -f -o gpurun_out/prof_gemmtc_r02 $CMD > gpurun_out/ncu_gtc.log 2>&1
tail -2 gpurun_out/ncu_gtc.log
