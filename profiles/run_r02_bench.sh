# round 2: the measurement set behind DESIGN.md section 5 (one B200)
mkdir -p gpurun_out
python bench.py > gpurun_out/r02_bench_default.log 2>&1
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r02_bench_reference.log 2>&1
python bench.py --strehl --steps 10 --no-cpu-baseline --no-variants > gpurun_out/r02_bench_strehl.log 2>&1
python bench.py --denoise --steps 3 --no-cpu-baseline --no-variants > gpurun_out/r02_bench_denoise.log 2>&1
python bench.py --workload 10x10 --steps 50 --no-cpu-baseline --no-variants > gpurun_out/r02_bench_10x10.log 2>&1
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-variants"
$CMD > gpurun_out/plain_list.log 2>&1 && \
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_E4096.csv $CMD > gpurun_out/ncu_list.log 2>&1
tail -c 300 gpurun_out/r02_bench_default.log
