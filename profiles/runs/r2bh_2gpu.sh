# 2 GPUs, final code: the driver's commands for N = 2 (own arm and reference arm)
set -x
mkdir -p gpurun_out
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 5 --warmup 3 ) > gpurun_out/c54_bench_2gpu.json 2> gpurun_out/c54_bench_2gpu.err
tail -3 gpurun_out/c54_bench_2gpu.err
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 ) > gpurun_out/c54_bench_2gpu_ref.json 2> gpurun_out/c54_bench_2gpu_ref.err
tail -3 gpurun_out/c54_bench_2gpu_ref.err
