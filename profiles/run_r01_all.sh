#!/bin/bash
# whole GPU suite, smoke, default bench, then the optional-path benches
cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/gpu_tests.log
tail -6 gpurun_out/gpu_tests.log | cut -c1-250
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_final.log 2>&1
tail -1 gpurun_out/bench_final.log | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('default', d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['cpu_baseline']['value'])"
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --denoise > gpurun_out/bench_denoise.log 2>&1
tail -1 gpurun_out/bench_denoise.log | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('denoise', d['value'], d['ms_per_step'])"
