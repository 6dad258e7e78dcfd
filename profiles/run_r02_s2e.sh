mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-variants  > gpurun_out/s2_b_ws.log 2>&1
python - <<'PY'
import json
d=json.loads(open('gpurun_out/s2_b_ws.log').read().strip().split('\n')[-1])
print(d['value'], d['ms_per_step'], d['roofline']['ms_per_launch'], d['e2e']['value'])
PY
CMD="python bench.py --steps 2 --warmup 3 --envs 1024 --no-cpu-baseline --no-variants "
ncu --profile-from-start off --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.per_cycle_active --clock-control none -k regex:wfs_frame -c 2 --csv --log-file gpurun_out/s2_inst_ws.csv $CMD > gpurun_out/s2_ncu.log 2>&1
grep -E "wfs_frame" gpurun_out/s2_inst_ws.csv | cut -d, -f5,13- | head -8
