set -x
mkdir -p gpurun_out
( time timeout 1200 python bench.py > gpurun_out/c53_bench_default.json 2> gpurun_out/c53_bench_default.err ) 2> gpurun_out/c53_time.txt
tail -2 gpurun_out/c53_bench_default.err; cat gpurun_out/c53_time.txt
