set -x
mkdir -p gpurun_out
( time timeout 1200 python bench.py > gpurun_out/c27_bench_default.json 2> gpurun_out/c27_bench_default.err ) 2> gpurun_out/c27_time.txt
tail -3 gpurun_out/c27_bench_default.err; cat gpurun_out/c27_time.txt
( time timeout 600 python bench.py --impl reference > gpurun_out/c27_bench_reference.json 2> gpurun_out/c27_bench_reference.err ) 2>> gpurun_out/c27_time.txt
tail -4 gpurun_out/c27_time.txt
