# first-error ordering (which exception class a stream with several defects raises): tests + batched / single campaigns
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/c40_tests.log 2>&1; tail -6 gpurun_out/c40_tests.log
( time timeout 1500 python tests/campaigns/fuzz_batches.py 120 96 9001 ) > gpurun_out/c40_fuzz_batches.txt 2>&1; tail -8 gpurun_out/c40_fuzz_batches.txt
( time timeout 1200 python tests/campaigns/fuzz_campaign.py 2000 71001 ) > gpurun_out/c40_fuzz_a.txt 2>&1; grep -E "trial|streams," gpurun_out/c40_fuzz_a.txt | tail -5
ls gpurun_out | wc -l
