set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/c7_topo.txt 2>&1
nproc >> gpurun_out/c7_topo.txt; free -g >> gpurun_out/c7_topo.txt
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 ) > gpurun_out/c7_bench_2gpu.json 2> gpurun_out/c7_bench_2gpu.err
tail -5 gpurun_out/c7_bench_2gpu.err
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 ) > gpurun_out/c7_bench_ref_2gpu.json 2> gpurun_out/c7_bench_ref_2gpu.err
tail -3 gpurun_out/c7_bench_ref_2gpu.err
cat gpurun_out/c7_bench_2gpu.json | head -c 600
