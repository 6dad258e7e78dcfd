# round 2, four B200s: the env-sharded bench of the final build
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02f_4gpu_devices.txt 2>&1
NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r02f_bench_4gpu.log 2>&1
tail -c 300 gpurun_out/r02f_bench_4gpu.log
