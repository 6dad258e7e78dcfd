mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/s2_t.log
tail -4 gpurun_out/s2_t.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-variants > gpurun_out/s2_b.log 2>&1
python - <<'PY'
import json
d=json.loads(open('gpurun_out/s2_b.log').read().strip().split('\n')[-1])
print(d['value'], d['ms_per_step'], d['roofline']['ms_per_launch'], d['e2e']['value'])
PY
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-variants"
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s2_launches_E4096.csv $CMD > gpurun_out/s2_ncu_list.log 2>&1
