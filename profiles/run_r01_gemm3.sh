#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gemm or closed_loop_against or extrusion" 2>&1 | tail -2
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('bench', d['value'], d['ms_per_step'], d['e2e']['value'])"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:gemm_tc -c 60 --csv --log-file gpurun_out/launches_gemm3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows=list(csv.reader(open('gpurun_out/launches_gemm3.csv')))
h=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
H=rows[h]; ik=H.index("Kernel Name"); iv=H.index("Metric Value"); ig=H.index("Grid Size")
d=collections.defaultdict(list)
for r in rows[h+1:]:
    if len(r)>iv: d[(r[ik][:40], r[ig])].append(float(r[iv].replace(',',''))/1e3)
for k,v in d.items(): print(k, len(v), "median %.1f us"%sorted(v)[len(v)//2])
PY
