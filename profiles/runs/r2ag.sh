# final captures, part 2: the self-synchronising chain and the encoder kernels (ncu --set full, no source import)
set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none -k regex:'jb_k1b_|jb_k1_huff_flat' -s 0 -c 12 -o gpurun_out/c31_prof_nr python bench.py --workload norestart --distinct 16 --steps 1 --warmup 1 --e2e-batch 32 --cpu-seconds 1 > gpurun_out/c31_ncu_nr.log 2>&1
tail -1 gpurun_out/c31_ncu_nr.log | cut -c1-120
ls -la gpurun_out
