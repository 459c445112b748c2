set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/c10_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/c10_tests.log
tail -4 gpurun_out/c10_tests.log
JB_K4_CLASSIC=1 timeout 600 python bench.py --workload encode --cpu-seconds 1 --steps 3 > gpurun_out/c10_bench_enc_classic.json 2> gpurun_out/c10_bench_enc_classic.err
timeout 600 python bench.py --workload encode --cpu-seconds 1 --steps 3 > gpurun_out/c10_bench_enc_fused.json 2> gpurun_out/c10_bench_enc_fused.err
tail -3 gpurun_out/c10_bench_enc_fused.err
( time timeout 900 python bench.py ) > gpurun_out/c10_bench_all.json 2> gpurun_out/c10_bench_all.err
tail -4 gpurun_out/c10_bench_all.err
