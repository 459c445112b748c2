# K1b un-stuff copy staged through shared memory, A/B
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_decode.py tests/test_gpu_fuzz.py tests/test_gpu_stress.py -m gpu -q -x > gpurun_out/c30_tests.log 2>&1; tail -3 gpurun_out/c30_tests.log
for v in directcopy "" directcopy ""; do
  if [ -n "$v" ]; then export JB_LIBDIR=$PWD/jpeglibrary_b200/lib_$v; else unset JB_LIBDIR; fi
  timeout 600 python bench.py --workload norestart --steps 5 --warmup 3 --e2e-batch 32 --cpu-seconds 1 --distinct 32 > gpurun_out/c30_bench_$v.json 2> gpurun_out/c30_bench_$v.err
  tail -2 gpurun_out/c30_bench_$v.err
  python - "$v" <<'PY'
import json,sys
for l in open('gpurun_out/c30_bench_%s.json'%sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print('VARIANT',sys.argv[1] or 'staged', d['ms_per_step'], d['roofline']['kernel_ms'], d['config']['parity']['max_abs_rgb_diff_vs_oracle'])
PY
done
unset JB_LIBDIR
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/c30_launches.csv python bench.py --workload norestart --steps 1 --warmup 1 --e2e-batch 32 --cpu-seconds 1 --distinct 32 > gpurun_out/c30_launches.log 2>&1
grep -E "jb_k1b_copy|jb_k1b_count" gpurun_out/c30_launches.csv | tail -4 | cut -c1-200
