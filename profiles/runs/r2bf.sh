# final state of round 2: smoke, GPU tests, the driver's default bench command and the reference arm
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/c52_tests.log 2>&1; tail -3 gpurun_out/c52_tests.log
( time timeout 1200 python bench.py > gpurun_out/c52_bench_default.json 2> gpurun_out/c52_bench_default.err ) 2> gpurun_out/c52_time.txt
tail -2 gpurun_out/c52_bench_default.err; cat gpurun_out/c52_time.txt
timeout 600 python bench.py --impl reference > gpurun_out/c52_bench_reference.json 2> gpurun_out/c52_bench_reference.err
