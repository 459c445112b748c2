set -x
mkdir -p gpurun_out
timeout 600 python profiles/two_context_overlap.py 1024 32 5 > gpurun_out/c28_overlap.txt 2>&1
tail -6 gpurun_out/c28_overlap.txt
