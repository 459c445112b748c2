set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'jb_k1c_progressive' -s 1 -c 1 -o gpurun_out/c21_prog python bench.py --workload progressive --batch 256 --distinct 8 --steps 1 --warmup 1 --e2e-batch 32 --cpu-seconds 1 > gpurun_out/c21_ncu.log 2>&1
tail -3 gpurun_out/c21_ncu.log | cut -c1-300
