mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu9.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu9.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_i_full.log 2>&1; echo "rc=$?"; grep -o '"value": [0-9.]*, "unit": "env-steps/s", "n_gpus"\|"ms_per_launch": [0-9.]*\|"ms_per_step": [0-9.]*' gpurun_out/bench_i_full.log
CMD="python bench.py --steps 2 --warmup 3 --envs 4096 --no-cpu-baseline"
$CMD > gpurun_out/plain_i.log 2>&1 && ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01_i.csv $CMD > gpurun_out/ncu_list_i.log 2>&1
