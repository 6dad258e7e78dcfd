# round 2, two B200s: the NCCL gradient all-reduce test of the batched SAC learner and the 2-GPU bench (config 5 section included)
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02f_2gpu_devices.txt 2>&1
timeout 900 python -m pytest tests/test_sac_learner.py tests/test_gpu_parity40.py -m gpu -q -k "nccl or per_environment_r0" > gpurun_out/r02f_nccl_test.log 2>&1
tail -3 gpurun_out/r02f_nccl_test.log
NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02f_bench_2gpu.log 2>&1
tail -c 400 gpurun_out/r02f_bench_2gpu.log
