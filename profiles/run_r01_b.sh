set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu2.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu2.log
for path in tensor tensor_fast simt; do
  timeout 300 python bench.py --envs 1024 --steps 5 --warmup 3 --no-cpu-baseline --wfs-path $path > gpurun_out/bench_1024_$path.log 2>&1; echo "rc=$?"; tail -c 1500 gpurun_out/bench_1024_$path.log
done
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_full2.log 2>&1; echo "rc=$?"; tail -c 2500 gpurun_out/bench_full2.log
