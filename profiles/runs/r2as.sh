set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/c42_tests.log 2>&1; tail -4 gpurun_out/c42_tests.log
( time timeout 1200 python bench.py > gpurun_out/c42_bench_default.json 2> gpurun_out/c42_bench_default.err ) 2> gpurun_out/c42_time.txt
tail -3 gpurun_out/c42_bench_default.err; cat gpurun_out/c42_time.txt
