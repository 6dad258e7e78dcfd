mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_l_full.log 2>&1; echo "rc=$?"; grep -o '"value": [0-9.]*, "unit": "env-steps/s"\|"ms_per_launch": [0-9.]*\|"ms_per_step": [0-9.]*' gpurun_out/bench_l_full.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --workload 10x10 > gpurun_out/bench_l_10x10.log 2>&1; echo "rc=$?"; grep -o '"value": [0-9.]*, "unit": "env-steps/s"\|"ms_per_launch": [0-9.]*\|"ms_per_step": [0-9.]*' gpurun_out/bench_l_10x10.log
