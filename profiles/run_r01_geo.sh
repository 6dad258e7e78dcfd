#!/bin/bash
# geometric controller / pupil sweep: parity tests, timings, then the step with and without it
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_geo.py tests/test_gpu_parity.py -m gpu -x -q -k "geo or sweep or trehl or single_environment" > gpurun_out/geo_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/geo_tests.log
tail -15 gpurun_out/geo_tests.log
timeout 300 python profiles/dev/geo_time.py 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --geo > gpurun_out/bench_geo.log 2>&1
tail -1 gpurun_out/bench_geo.log | cut -c1-330
