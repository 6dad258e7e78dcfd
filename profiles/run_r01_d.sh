set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu4.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu4.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_full4.log 2>&1; echo "rc=$?"; tail -c 1600 gpurun_out/bench_full4.log
CMD="python bench.py --steps 2 --warmup 3 --envs 4096 --no-cpu-baseline"
$CMD > gpurun_out/plain_d.log 2>&1 && ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01_d.csv $CMD > gpurun_out/ncu_list_d.log 2>&1
