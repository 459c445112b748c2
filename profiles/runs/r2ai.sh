# corrupted-stream campaign on the final code of round 2 (new seeds)
set -x
mkdir -p gpurun_out
( time timeout 900 python tests/campaigns/fuzz_campaign.py 1500 20261 ) > gpurun_out/c32_fuzz_a.txt 2>&1; tail -14 gpurun_out/c32_fuzz_a.txt
( time timeout 900 python tests/campaigns/fuzz_campaign.py 1000 20262 --more ) > gpurun_out/c32_fuzz_b.txt 2>&1; tail -14 gpurun_out/c32_fuzz_b.txt
