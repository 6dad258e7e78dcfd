# round 2: the measurement set behind DESIGN.md section 5 (one B200)
mkdir -p gpurun_out
python bench.py > gpurun_out/r02_bench_default.log 2>&1
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r02_bench_reference.log 2>&1
python bench.py --strehl --steps 10 --no-cpu-baseline --no-variants > gpurun_out/r02_bench_strehl.log 2>&1
python bench.py --strehl --strehl-peak --steps 10 --no-cpu-baseline --no-variants > gpurun_out/r02_bench_strehl_peak.log 2>&1
python bench.py --denoise --steps 3 --no-cpu-baseline --no-variants > gpurun_out/r02_bench_denoise.log 2>&1
python bench.py --workload 10x10 --steps 50 --no-cpu-baseline --no-variants > gpurun_out/r02_bench_10x10.log 2>&1
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-variants"
$CMD > gpurun_out/plain_list.log 2>&1 && \
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_E4096.csv $CMD > gpurun_out/ncu_list.log 2>&1
CMD="python bench.py --steps 2 --warmup 3 --envs 1024 --no-cpu-baseline --no-variants"
$CMD > gpurun_out/plain_umma.log 2>&1 && \
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:wfs_frame_umma -c 1 -o gpurun_out/prof_wfs_umma_r02_v4 $CMD > gpurun_out/ncu_umma.log 2>&1
for f in default reference strehl strehl_peak denoise 10x10; do python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_$f.log').read().strip().split('\n')[-1])
print('$f', d.get('value'), d.get('ms_per_step'), (d.get('e2e') or {}).get('value'))
PY
done
