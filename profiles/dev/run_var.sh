for v in 0 1 2 3; do
AOM_WFS_VARIANT=$v timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/var_$v.log 2>&1
echo "variant=$v rc=$?"; grep -o '"ms_per_launch": [0-9.]*\|"ms_per_step": [0-9.]*\|Error.*' gpurun_out/var_$v.log | head -3
done
AOM_WFS_VARIANT=1 timeout 600 python -m pytest tests -m gpu -q -x -k "wfs or 40x40 or frame" 2>&1 | tail -3
