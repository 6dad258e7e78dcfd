for ns in 0 1000 2000 3000 4000; do
d=$((ns*16))
timeout 300 python bench.py --steps 5 --warmup 3 --envs 1024 --no-cpu-baseline --wfs-dbg $d > gpurun_out/dbg_s$ns.log 2>&1
echo "stagger=$ns rc=$?"; tail -c 1600 gpurun_out/dbg_s$ns.log | grep -o '"ms_per_launch": [0-9.]*\|Error.*' 
done
