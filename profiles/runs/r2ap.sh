set -x
mkdir -p gpurun_out
( time timeout 1500 python tests/campaigns/fuzz_batches.py 120 96 9001 ) > gpurun_out/c39_fuzz_batches.txt 2>&1; tail -12 gpurun_out/c39_fuzz_batches.txt
ls gpurun_out | wc -l
