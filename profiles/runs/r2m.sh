set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/c13_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/c13_tests.log
tail -4 gpurun_out/c13_tests.log
timeout 600 python bench.py --workload restart --distinct 128 --cpu-seconds 1 --e2e-batch 64 > gpurun_out/c13_bench.json 2> gpurun_out/c13_bench.err
timeout 600 python bench.py --workload norestart --distinct 32 --cpu-seconds 1 --steps 3 > gpurun_out/c13_bench_nr.json 2> gpurun_out/c13_bench_nr.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'jb_k1b_sync' -s 0 -c 2 -o gpurun_out/c13_prof_nr python bench.py --workload norestart --distinct 16 --steps 1 --warmup 1 --cpu-seconds 1 > gpurun_out/c13_ncu.log 2>&1
tail -2 gpurun_out/c13_ncu.log | cut -c1-200
