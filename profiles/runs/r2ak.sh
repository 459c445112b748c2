# corrupted-stream campaign over layouts and scan scripts the other sets do not hold
set -x
mkdir -p gpurun_out
( time timeout 1500 python tests/campaigns/fuzz_campaign.py 3000 41001 --wide ) > gpurun_out/c34_fuzz_wide.txt 2>&1; tail -30 gpurun_out/c34_fuzz_wide.txt
ls gpurun_out | wc -l
