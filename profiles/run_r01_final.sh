mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_final.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_final.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_final.log
( time timeout 900 python bench.py ) > gpurun_out/bench_final.log 2>&1; echo "bench rc=$?"; grep -o '"value": [0-9.]*, "unit": "env-steps/s"\|"ms_per_step": [0-9.]*\|"frac": [0-9.]*\|"share_of_step": [0-9.]*\|real.*' gpurun_out/bench_final.log
( time timeout 900 python bench.py --impl reference --steps 20 --warmup 3 ) > gpurun_out/bench_final_ref.log 2>&1; echo "ref rc=$?"; grep -o '"value": [0-9.]*, "unit": "env-steps/s", "n_gpus"' gpurun_out/bench_final_ref.log
