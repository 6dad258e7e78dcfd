mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "seam or generations" 2>&1 | tail -4 > gpurun_out/s2_t.log
tail -2 gpurun_out/s2_t.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-variants > gpurun_out/s2_b.log 2>&1
python - <<'PY'
import json
d=json.loads(open('gpurun_out/s2_b.log').read().strip().split('\n')[-1])
print(d['value'], d['ms_per_step'], d['roofline']['ms_per_launch'], d['e2e']['value'])
PY
