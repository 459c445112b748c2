set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/c14_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/c14_tests.log
tail -4 gpurun_out/c14_tests.log
( time timeout 900 python bench.py ) > gpurun_out/c14_bench_all.json 2> gpurun_out/c14_bench_all.err
tail -4 gpurun_out/c14_bench_all.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'jb_k0_restart_scan|jb_k1_huff_flat|jb_k2_idct_color_warp' -s 3 -c 3 -o gpurun_out/c14_prof python bench.py --workload restart --distinct 16 --steps 1 --warmup 2 --e2e-batch 32 --cpu-seconds 1 > gpurun_out/c14_ncu.log 2>&1
tail -2 gpurun_out/c14_ncu.log | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 10 -c 40 --csv --log-file gpurun_out/c14_launches.csv python bench.py --workload restart --distinct 16 --steps 2 --warmup 1 --e2e-batch 32 --cpu-seconds 1 > gpurun_out/c14_launches.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:'jb_k1b_|jb_k1_huff_flat' -s 0 -c 12 -o gpurun_out/c14_prof_nr python bench.py --workload norestart --distinct 16 --steps 1 --warmup 1 --cpu-seconds 1 > gpurun_out/c14_ncu_nr.log 2>&1
bash profiles/runs/r2_sanitizers.sh
