# round 2, final build: the measurement set behind DESIGN.md section 5 (one B200)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02f_smoke.log 2>&1; tail -2 gpurun_out/r02f_smoke.log
python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r02f_gpu_tests.log; tail -2 gpurun_out/r02f_gpu_tests.log
python bench.py > gpurun_out/r02f_bench_default.log 2>&1
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r02f_bench_reference.log 2>&1
python bench.py --wfs-path umma --steps 10 --no-cpu-baseline --no-variants > gpurun_out/r02f_bench_fused_sensor_kernel.log 2>&1
python bench.py --strehl --steps 10 --no-cpu-baseline --no-variants > gpurun_out/r02f_bench_strehl.log 2>&1
python bench.py --denoise --steps 3 --no-cpu-baseline --no-variants > gpurun_out/r02f_bench_denoise.log 2>&1
python bench.py --workload 10x10 --steps 50 --no-cpu-baseline --no-variants > gpurun_out/r02f_bench_10x10.log 2>&1
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-variants"
$CMD > gpurun_out/plain_list.log 2>&1 && \
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02f_launches_E4096.csv $CMD > gpurun_out/ncu_list.log 2>&1
for f in default reference fused_sensor_kernel strehl denoise 10x10; do python - <<PY
import json
d=json.loads(open('gpurun_out/r02f_bench_$f.log').read().strip().split('\n')[-1])
print('$f', d.get('value'), d.get('ms_per_step'), (d.get('e2e') or {}).get('value'), (d.get('roofline') or {}).get('ms_per_launch'))
PY
done
