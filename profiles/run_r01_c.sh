set -x
mkdir -p gpurun_out
timeout 180 python -m pytest tests/test_gpu_parity.py -k "gemm" -x -q > gpurun_out/pytest_gemm.log 2>&1; rc=$?; echo "gemm pytest rc=$rc"; tail -30 gpurun_out/pytest_gemm.log
nvidia-smi --query-gpu=name,clocks.sm,memory.used --format=csv
if [ $rc -ne 0 ]; then export AOM_GEMM_PATH=simt; echo "FALLING BACK TO SIMT GEMM FOR THE REST"; fi
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu3.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu3.log
timeout 300 python bench.py --envs 1024 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_1024_c.log 2>&1; echo "rc=$?"; tail -c 1600 gpurun_out/bench_1024_c.log
AOM_GEMM_PATH=simt timeout 300 python bench.py --envs 1024 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_1024_c_simtgemm.log 2>&1; echo "rc=$?"; tail -c 1600 gpurun_out/bench_1024_c_simtgemm.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_full3.log 2>&1; echo "rc=$?"; tail -c 1600 gpurun_out/bench_full3.log
