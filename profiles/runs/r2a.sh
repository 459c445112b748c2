set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c1_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/c1_tests.log
tail -5 gpurun_out/c1_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/c1_bench_packed.json 2> gpurun_out/c1_bench_packed.err
JB_LIBDIR=/root/repo/jpeglibrary_b200/lib_scalar timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/c1_bench_scalar.json 2> gpurun_out/c1_bench_scalar.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'jb_k1_huff_flat|jb_k2_idct_color_warp' -s 4 -c 2 -o gpurun_out/c1_prof python bench.py --steps 1 --warmup 1 --batch 1024 --e2e-batch 32 --cpu-images 16 > gpurun_out/c1_ncu.log 2>&1
tail -3 gpurun_out/c1_ncu.log
cat gpurun_out/c1_bench_packed.json | head -c 1500
