# round 2, last build: refresh of the default line, the launch list and the smoke / test logs
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02f_smoke.log 2>&1; tail -1 gpurun_out/r02f_smoke.log
python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r02f_gpu_tests.log; tail -1 gpurun_out/r02f_gpu_tests.log
python bench.py > gpurun_out/r02f_bench_default.log 2>&1
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-variants"
$CMD > gpurun_out/plain_list.log 2>&1 && \
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02f_launches_E4096.csv $CMD > gpurun_out/ncu_list.log 2>&1
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02f_bench_default.log').read().strip().split('\n')[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['ms_per_launch'], d['roofline']['frac'], d['reset']['ms'])
PY
