# final captures of round 2: ncu --set full of the headline kernels, the self-synchronising chain, the encoder;
# launch list of the default-workload command; compute-sanitizer racecheck / synccheck / memcheck
set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'jb_k0_restart_scan|jb_k1_huff_flat|jb_k2_idct_color_warp' -s 3 -c 3 -o gpurun_out/c31_prof python bench.py --workload restart --distinct 16 --steps 1 --warmup 2 --e2e-batch 32 --cpu-seconds 1 > gpurun_out/c31_ncu.log 2>&1
tail -1 gpurun_out/c31_ncu.log | cut -c1-120
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 10 -c 40 --csv --log-file gpurun_out/c31_launches.csv python bench.py --workload restart --distinct 16 --steps 2 --warmup 1 --e2e-batch 32 --cpu-seconds 1 > gpurun_out/c31_launches.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:'jb_k1b_|jb_k1_huff_flat' -s 0 -c 12 -o gpurun_out/c31_prof_nr python bench.py --workload norestart --distinct 16 --steps 1 --warmup 1 --e2e-batch 32 --cpu-seconds 1 > gpurun_out/c31_ncu_nr.log 2>&1
tail -1 gpurun_out/c31_ncu_nr.log | cut -c1-120
timeout 900 ncu --set full --clock-control none -k regex:'jb_k3|jb_k4' -s 0 -c 8 -o gpurun_out/c31_prof_enc python bench.py --workload encode --batch 128 --distinct 16 --steps 1 --warmup 1 --cpu-seconds 1 > gpurun_out/c31_ncu_enc.log 2>&1
tail -1 gpurun_out/c31_ncu_enc.log | cut -c1-120
bash profiles/runs/r2_sanitizers.sh
