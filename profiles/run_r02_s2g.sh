mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-variants"
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"oz_gather_slice|extrude_scatter|oz_gemm_kernel" -s 20 -c 6 -f -o gpurun_out/prof_extrude_r02 $CMD > gpurun_out/ncu_ex.log 2>&1
tail -2 gpurun_out/ncu_ex.log
