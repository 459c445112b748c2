set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/c46_tests.log 2>&1; tail -5 gpurun_out/c46_tests.log
( time timeout 1500 python tests/campaigns/fuzz_shapes.py 1000 4245 ) > gpurun_out/c46_shapes.txt 2>&1; tail -12 gpurun_out/c46_shapes.txt
ls gpurun_out | wc -l
