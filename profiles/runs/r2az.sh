set -x
mkdir -p gpurun_out
( time timeout 1500 python tests/campaigns/fuzz_shapes.py 1200 4246 ) > gpurun_out/c47_shapes.txt 2>&1; tail -8 gpurun_out/c47_shapes.txt
( time timeout 1500 python tests/campaigns/fuzz_batches.py 200 128 9003 ) > gpurun_out/c47_fuzz_batches.txt 2>&1; tail -4 gpurun_out/c47_fuzz_batches.txt
ls gpurun_out | wc -l
