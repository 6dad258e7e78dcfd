#!/bin/bash
# does the two-MMA stage-2 variant (tensor_fast) hold the 1e-4 parity bar, and what does it buy?
cd /root/repo
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --wfs-path tensor_fast > gpurun_out/bench_fast.log 2>&1
tail -1 gpurun_out/bench_fast.log | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('fast', d['value'], d['ms_per_step'], d['roofline']['ms_per_launch'])"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_full.log 2>&1
tail -1 gpurun_out/bench_full.log | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('full', d['value'], d['ms_per_step'], d['roofline']['ms_per_launch'])"
AOM_WFS_PATH=tensor_fast timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q > gpurun_out/fast_tests.log 2>&1
tail -25 gpurun_out/fast_tests.log | cut -c1-200
