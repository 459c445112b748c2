set -x
mkdir -p gpurun_out
( time timeout 900 python tests/campaigns/fuzz_campaign.py 2500 81001 --cmyk ) > gpurun_out/c51_fuzz_cmyk.txt 2>&1; grep -E "trial|identical|streams," gpurun_out/c51_fuzz_cmyk.txt | tail -14
( time timeout 900 python tests/campaigns/fuzz_campaign.py 2500 81002 --wide ) > gpurun_out/c51_fuzz_wide.txt 2>&1; grep -E "trial|streams," gpurun_out/c51_fuzz_wide.txt | tail -6
ls gpurun_out | wc -l
