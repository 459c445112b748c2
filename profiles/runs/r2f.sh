set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/c6_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/c6_tests.log
tail -4 gpurun_out/c6_tests.log
for v in lib lib_coop lib_s208; do
  JB_LIBDIR=/root/repo/jpeglibrary_b200/$v timeout 600 python bench.py --workload restart --distinct 128 --cpu-seconds 1 > gpurun_out/c6_bench_$v.json 2> gpurun_out/c6_bench_$v.err
  JB_LIBDIR=/root/repo/jpeglibrary_b200/$v timeout 600 python bench.py --workload norestart --distinct 32 --cpu-seconds 1 --steps 3 > gpurun_out/c6_bench_nr_$v.json 2> gpurun_out/c6_bench_nr_$v.err
done
JB_LIBDIR=/root/repo/jpeglibrary_b200/lib_s208 timeout 600 python -m pytest tests -m gpu -q -x -k "synthetic_streams or self_synchronising or fuzz or corrupted" > gpurun_out/c6_tests_s208.log 2>&1
tail -3 gpurun_out/c6_tests_s208.log
