#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_denoiser.py tests/test_gpu_parity.py -m gpu -x -q -k "denois or fused or in_place" > gpurun_out/dn_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/dn_tests.log
tail -12 gpurun_out/dn_tests.log | cut -c1-250
timeout 300 python profiles/dev/denoiser_time.py 2>&1 | tail -5
