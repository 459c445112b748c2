set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -k "progressive or fuzz or corrupted or several_scans or mixed" > gpurun_out/c8_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/c8_tests.log
tail -8 gpurun_out/c8_tests.log
for g in 4 2 1; do
  JB_K1C_GROUPS=$g timeout 600 python bench.py --workload progressive --distinct 32 --cpu-seconds 1 --steps 3 > gpurun_out/c8_bench_prog_g$g.json 2> gpurun_out/c8_bench_prog_g$g.err
  tail -2 gpurun_out/c8_bench_prog_g$g.err
done
