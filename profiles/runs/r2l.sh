set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/c12_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/c12_tests.log
tail -4 gpurun_out/c12_tests.log
for v in lib lib_k0ldg; do
  JB_LIBDIR=/root/repo/jpeglibrary_b200/$v timeout 600 python bench.py --workload restart --distinct 128 --cpu-seconds 1 --e2e-batch 64 > gpurun_out/c12_bench_$v.json 2> gpurun_out/c12_bench_$v.err
  JB_LIBDIR=/root/repo/jpeglibrary_b200/$v timeout 600 python bench.py --workload progressive --distinct 32 --cpu-seconds 1 --steps 3 > gpurun_out/c12_bench_pr_$v.json 2> gpurun_out/c12_bench_pr_$v.err
done
timeout 300 python profiles/small_batch.py > gpurun_out/c12_small_batch.log 2>&1
tail -12 gpurun_out/c12_small_batch.log
