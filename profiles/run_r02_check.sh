mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/t3.log
tail -4 gpurun_out/t3.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-variants > gpurun_out/b3.log 2>&1
python - <<'PY'
import json
d=json.loads(open('gpurun_out/b3.log').read().strip().split('\n')[-1])
print(d['value'], d['ms_per_step'], d['roofline']['ms_per_launch'], d['e2e']['value'])
PY
