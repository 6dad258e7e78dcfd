mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/t3.log
tail -4 gpurun_out/t3.log
CMD="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-variants"
$CMD > gpurun_out/b3.log 2>&1
python - <<'PY'
import json
d=json.loads(open('gpurun_out/b3.log').read().strip().split('\n')[-1])
print(d['value'], d['ms_per_step'], d['roofline']['ms_per_launch'], d['e2e']['value'])
PY
CMD2="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-variants"
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches3.csv $CMD2 > gpurun_out/ncu_list3.log 2>&1
python bench.py --denoise --steps 3 --no-cpu-baseline --no-variants > gpurun_out/b3_denoise.log 2>&1
tail -c 400 gpurun_out/b3_denoise.log | head -c 200
