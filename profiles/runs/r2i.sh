set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/c9_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/c9_tests.log
tail -4 gpurun_out/c9_tests.log
for v in lib lib_k1b1; do
  JB_LIBDIR=/root/repo/jpeglibrary_b200/$v timeout 600 python bench.py --workload norestart --distinct 32 --cpu-seconds 1 --steps 3 > gpurun_out/c9_bench_nr_$v.json 2> gpurun_out/c9_bench_nr_$v.err
done
for cfg in "32 256" "64 256" "16 256" "32 512" "64 512"; do
  set -- $cfg
  timeout 600 python bench.py --workload restart --distinct 16 --cpu-seconds 1 --steps 5 --e2e-chunk $1 --e2e-batch $2 > gpurun_out/c9_e2e_c$1_b$2.json 2> gpurun_out/c9_e2e_c$1_b$2.err
done
