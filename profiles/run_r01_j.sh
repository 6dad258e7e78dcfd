mkdir -p gpurun_out
python profiles/dev/wfs_probe.py atmos 2>&1 | tail -5
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu10.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu10.log
for path in tensor tensor_staged; do
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --wfs-path $path > gpurun_out/bench_j_$path.log 2>&1; echo "rc=$?"; grep -o '"value": [0-9.]*, "unit": "env-steps/s", "n_gpus"\|"ms_per_launch": [0-9.]*\|"ms_per_step": [0-9.]*' gpurun_out/bench_j_$path.log
done
