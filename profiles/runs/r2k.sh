set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/c11_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/c11_tests.log
tail -4 gpurun_out/c11_tests.log
for v in lib lib_groupcopy; do
  JB_LIBDIR=/root/repo/jpeglibrary_b200/$v timeout 600 python bench.py --workload restart --distinct 128 --cpu-seconds 1 --e2e-batch 64 > gpurun_out/c11_bench_$v.json 2> gpurun_out/c11_bench_$v.err
  JB_LIBDIR=/root/repo/jpeglibrary_b200/$v timeout 600 python bench.py --workload norestart --distinct 32 --cpu-seconds 1 --steps 3 > gpurun_out/c11_bench_nr_$v.json 2> gpurun_out/c11_bench_nr_$v.err
done
