mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu6.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu6.log
timeout 600 python bench.py --steps 5 --warmup 3 --envs 1024 --no-cpu-baseline > gpurun_out/bench_f_1024.log 2>&1; echo "rc=$?"; grep -o '"value": [0-9.]*, "unit": "env-steps/s", "n_gpus"\|"ms_per_launch": [0-9.]*\|"ms_per_step": [0-9.]*' gpurun_out/bench_f_1024.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_f_full.log 2>&1; echo "rc=$?"; grep -o '"value": [0-9.]*, "unit": "env-steps/s", "n_gpus"\|"ms_per_launch": [0-9.]*\|"ms_per_step": [0-9.]*' gpurun_out/bench_f_full.log
