set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/c3_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/c3_tests.log
tail -6 gpurun_out/c3_tests.log
for v in lib_k1s1 lib lib_k1s3; do
  JB_LIBDIR=/root/repo/jpeglibrary_b200/$v timeout 600 python bench.py --workload restart --distinct 128 --cpu-seconds 1 > gpurun_out/c3_bench_$v.json 2> gpurun_out/c3_bench_$v.err
  JB_LIBDIR=/root/repo/jpeglibrary_b200/$v timeout 600 python bench.py --workload norestart --distinct 32 --cpu-seconds 1 --steps 3 > gpurun_out/c3_bench_nr_$v.json 2> gpurun_out/c3_bench_nr_$v.err
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'jb_k1_huff_flat' -s 1 -c 1 -o gpurun_out/c3_prof python bench.py --workload restart --distinct 16 --steps 1 --warmup 2 --e2e-batch 32 --cpu-seconds 1 > gpurun_out/c3_ncu.log 2>&1
tail -2 gpurun_out/c3_ncu.log | cut -c1-200
