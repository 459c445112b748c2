# per-component MCU limits for sequential scan-list frames (EOI at a restart boundary): tests + campaigns
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/c37_tests.log 2>&1; tail -4 gpurun_out/c37_tests.log
( time timeout 1200 python tests/campaigns/fuzz_campaign.py 3000 61002 --more ) > gpurun_out/c37_fuzz_b.txt 2>&1; grep -E "trial|streams,|identical" gpurun_out/c37_fuzz_b.txt | tail -14
( time timeout 1500 python tests/campaigns/fuzz_campaign.py 3000 61003 --wide ) > gpurun_out/c37_fuzz_c.txt 2>&1; grep -E "trial|streams,|identical" gpurun_out/c37_fuzz_c.txt | tail -14
ls gpurun_out | wc -l
