# 8 GPUs, final code of round 2 (batched H2D): the driver's command for N = 8
set -x
mkdir -p gpurun_out
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 ) > gpurun_out/c36_bench_8gpu.json 2> gpurun_out/c36_bench_8gpu.err
tail -5 gpurun_out/c36_bench_8gpu.err
cat gpurun_out/c36_bench_8gpu.json | head -c 2500
