set -x
mkdir -p gpurun_out
( time timeout 1500 python tests/campaigns/fuzz_shapes.py 240 4250 2200 ) > gpurun_out/c49_shapes_large.txt 2>&1; tail -8 gpurun_out/c49_shapes_large.txt
ls gpurun_out | wc -l
