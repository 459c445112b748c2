set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/c41_tests.log 2>&1; tail -5 gpurun_out/c41_tests.log
( time timeout 1500 python tests/campaigns/fuzz_batches.py 300 128 9002 ) > gpurun_out/c41_fuzz_batches.txt 2>&1; tail -6 gpurun_out/c41_fuzz_batches.txt
ls gpurun_out | wc -l
