# final captures, part 1: the headline kernels (ncu --set full) and the launch list of the same command
set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'jb_k0_restart_scan|jb_k1_huff_flat|jb_k2_idct_color_warp' -s 3 -c 3 -o gpurun_out/c31_prof python bench.py --workload restart --distinct 16 --steps 1 --warmup 2 --e2e-batch 32 --cpu-seconds 1 > gpurun_out/c31_ncu.log 2>&1
tail -1 gpurun_out/c31_ncu.log | cut -c1-120
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 10 -c 40 --csv --log-file gpurun_out/c31_launches.csv python bench.py --workload restart --distinct 16 --steps 2 --warmup 1 --e2e-batch 32 --cpu-seconds 1 > gpurun_out/c31_launches.log 2>&1
ls -la gpurun_out
