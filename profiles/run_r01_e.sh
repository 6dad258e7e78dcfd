set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu5.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/pytest_gpu5.log
for path in tensor tensor_reg; do
timeout 600 python bench.py --steps 5 --warmup 3 --envs 1024 --no-cpu-baseline --wfs-path $path > gpurun_out/bench_e_1024_$path.log 2>&1; echo "rc=$?"; tail -c 1500 gpurun_out/bench_e_1024_$path.log
done
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_e_full.log 2>&1; echo "rc=$?"; tail -c 1500 gpurun_out/bench_e_full.log
