mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu11.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu11.log
for path in tcgen05 tensor; do
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --wfs-path $path > gpurun_out/bench_k_$path.log 2>&1; echo "$path rc=$?"; grep -o '"value": [0-9.]*, "unit": "env-steps/s", "n_gpus"\|"ms_per_launch": [0-9.]*\|"ms_per_step": [0-9.]*' gpurun_out/bench_k_$path.log
done
