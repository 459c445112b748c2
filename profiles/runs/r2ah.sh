# final captures, part 3: the encoder kernels (ncu --set full)
set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none -k regex:'jb_k3|jb_k4' -s 0 -c 8 -o gpurun_out/c31_prof_enc python bench.py --workload encode --batch 128 --distinct 16 --steps 1 --warmup 1 --cpu-seconds 1 > gpurun_out/c31_ncu_enc.log 2>&1
tail -1 gpurun_out/c31_ncu_enc.log | cut -c1-120
ls -la gpurun_out
