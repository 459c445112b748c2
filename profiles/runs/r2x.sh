set -x
mkdir -p gpurun_out
timeout 300 python profiles/prog_trace.py 1024 8 > gpurun_out/c24_trace.txt 2>&1
tail -11 gpurun_out/c24_trace.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'jb_k1c_progressive' -s 1 -c 1 -o gpurun_out/c24_prog python bench.py --workload progressive --distinct 8 --steps 1 --warmup 1 --e2e-batch 32 --cpu-seconds 1 > gpurun_out/c24_ncu.log 2>&1
tail -2 gpurun_out/c24_ncu.log | cut -c1-200
