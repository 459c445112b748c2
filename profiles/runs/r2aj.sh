# scan-header fixes: tests, then a longer corrupted-stream campaign with new seeds
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fuzz.py tests/test_gpu_decode.py -m gpu -q -x > gpurun_out/c33_tests.log 2>&1; tail -3 gpurun_out/c33_tests.log
( time timeout 1200 python tests/campaigns/fuzz_campaign.py 4000 31001 ) > gpurun_out/c33_fuzz_a.txt 2>&1; tail -12 gpurun_out/c33_fuzz_a.txt
( time timeout 1200 python tests/campaigns/fuzz_campaign.py 4000 31002 --more ) > gpurun_out/c33_fuzz_b.txt 2>&1; tail -16 gpurun_out/c33_fuzz_b.txt
ls gpurun_out | wc -l
