set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/c2_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/c2_tests.log
tail -15 gpurun_out/c2_tests.log
( time timeout 900 python bench.py ) > gpurun_out/c2_bench_all.json 2> gpurun_out/c2_bench_all.err
tail -3 gpurun_out/c2_bench_all.err
JB_LIBDIR=/root/repo/jpeglibrary_b200/lib_scalar timeout 600 python bench.py --workload restart --distinct 16 --cpu-seconds 2 > gpurun_out/c2_bench_scalar.json 2> gpurun_out/c2_bench_scalar.err
timeout 600 python bench.py --workload restart --distinct 16 --cpu-seconds 2 > gpurun_out/c2_bench_packed.json 2> gpurun_out/c2_bench_packed.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'jb_k1_huff_flat|jb_k2_idct_color_warp' -s 2 -c 2 -o gpurun_out/c2_prof python bench.py --workload restart --distinct 16 --steps 1 --warmup 2 --e2e-batch 32 --cpu-seconds 1 > gpurun_out/c2_ncu.log 2>&1
tail -2 gpurun_out/c2_ncu.log | cut -c1-300
