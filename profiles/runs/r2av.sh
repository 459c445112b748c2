set -x
mkdir -p gpurun_out
( time timeout 1500 python tests/campaigns/fuzz_shapes.py 600 4242 ) > gpurun_out/c45_shapes.txt 2>&1; tail -25 gpurun_out/c45_shapes.txt
ls gpurun_out | wc -l
