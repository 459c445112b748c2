set -x
mkdir -p gpurun_out
timeout 300 python profiles/runs/dbg1.py > gpurun_out/c4_dbg1.log 2>&1
cat gpurun_out/c4_dbg1.log | tail -40
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/c4_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/c4_tests.log
tail -6 gpurun_out/c4_tests.log
for v in lib lib_k1s4; do
  JB_LIBDIR=/root/repo/jpeglibrary_b200/$v timeout 600 python bench.py --workload restart --distinct 128 --cpu-seconds 1 > gpurun_out/c4_bench_$v.json 2> gpurun_out/c4_bench_$v.err
done
bash profiles/runs/r2_sanitizers.sh
