mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu7.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/pytest_gpu7.log
