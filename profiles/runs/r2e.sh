set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/c5_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/c5_tests.log
tail -6 gpurun_out/c5_tests.log
( time timeout 900 python bench.py ) > gpurun_out/c5_bench_all.json 2> gpurun_out/c5_bench_all.err
tail -4 gpurun_out/c5_bench_all.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'jb_k1_huff_flat|jb_k2_idct_color_warp' -s 2 -c 2 -o gpurun_out/c5_prof python bench.py --workload restart --distinct 16 --steps 1 --warmup 2 --e2e-batch 32 --cpu-seconds 1 > gpurun_out/c5_ncu.log 2>&1
tail -2 gpurun_out/c5_ncu.log | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 60 --csv --log-file gpurun_out/c5_launches.csv python bench.py --workload restart --distinct 16 --steps 2 --warmup 1 --e2e-batch 32 --cpu-seconds 1 > gpurun_out/c5_launches.log 2>&1
