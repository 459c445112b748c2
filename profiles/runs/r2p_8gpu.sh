set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/c16_topo.txt 2>&1
nproc >> gpurun_out/c16_topo.txt; free -g >> gpurun_out/c16_topo.txt; lscpu | grep -i "numa\|model name\|socket" >> gpurun_out/c16_topo.txt
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 --workload restart ) > gpurun_out/c16_bench_8gpu.json 2> gpurun_out/c16_bench_8gpu.err
tail -5 gpurun_out/c16_bench_8gpu.err
cat gpurun_out/c16_bench_8gpu.json | head -c 3000
