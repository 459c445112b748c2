set -x
mkdir -p gpurun_out
( time timeout 600 python tests/campaigns/fuzz_campaign.py 2000 91001 ) > gpurun_out/c55_fuzz_a.txt 2>&1; grep -E "trial|streams," gpurun_out/c55_fuzz_a.txt | tail -4
( time timeout 600 python tests/campaigns/fuzz_campaign.py 1500 91002 --more ) > gpurun_out/c55_fuzz_b.txt 2>&1; grep -E "trial|streams," gpurun_out/c55_fuzz_b.txt | tail -4
( time timeout 600 python tests/campaigns/fuzz_shapes.py 800 4260 ) > gpurun_out/c55_shapes.txt 2>&1; tail -3 gpurun_out/c55_shapes.txt
