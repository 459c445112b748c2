# arena tail = FF (a trailing FF without a next byte is dropped like in the reference): full tests + campaigns, new seeds
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/c35_tests.log 2>&1; tail -4 gpurun_out/c35_tests.log
( time timeout 1200 python tests/campaigns/fuzz_campaign.py 3000 51001 ) > gpurun_out/c35_fuzz_a.txt 2>&1; tail -4 gpurun_out/c35_fuzz_a.txt
( time timeout 1200 python tests/campaigns/fuzz_campaign.py 3000 51002 --more ) > gpurun_out/c35_fuzz_b.txt 2>&1; grep -E "trial|streams," gpurun_out/c35_fuzz_b.txt | tail -8
( time timeout 1500 python tests/campaigns/fuzz_campaign.py 3000 51003 --wide ) > gpurun_out/c35_fuzz_c.txt 2>&1; grep -E "trial|streams," gpurun_out/c35_fuzz_c.txt | tail -8
ls gpurun_out | wc -l
