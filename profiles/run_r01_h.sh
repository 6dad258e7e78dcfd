mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu_h.txt; nproc >> gpurun_out/gpu_h.txt
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu8.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu8.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_h.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_h.log
( time timeout 900 python bench.py ) > gpurun_out/bench_h_default.log 2>&1; echo "bench rc=$?"; tail -c 2600 gpurun_out/bench_h_default.log
( time timeout 900 python bench.py --impl reference --steps 20 --warmup 3 ) > gpurun_out/bench_h_ref.log 2>&1; echo "ref rc=$?"; tail -c 1500 gpurun_out/bench_h_ref.log
