# K1 symbol step: magnitude by one funnel shift; slot address as IMAD (FMA pipe)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_decode.py tests/test_gpu_fuzz.py -m gpu -q -x > gpurun_out/c29_tests.log 2>&1; tail -3 gpurun_out/c29_tests.log
for v in base noimad "" base ""; do
  if [ -n "$v" ]; then export JB_LIBDIR=$PWD/jpeglibrary_b200/lib_$v; else unset JB_LIBDIR; fi
  timeout 600 python bench.py --workload restart --steps 5 --warmup 3 --e2e-batch 32 --cpu-seconds 1 --distinct 32 > gpurun_out/c29_bench_$v.json 2> gpurun_out/c29_bench_$v.err
  tail -2 gpurun_out/c29_bench_$v.err
  python - "$v" <<'PY'
import json,sys
for l in open('gpurun_out/c29_bench_%s.json'%sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print('VARIANT',sys.argv[1] or 'default', d['ms_per_step'], d['roofline']['kernel_ms'], d['config']['parity']['max_abs_rgb_diff_vs_oracle'])
PY
done
unset JB_LIBDIR
timeout 600 python bench.py --workload norestart --steps 5 --warmup 3 --e2e-batch 32 --cpu-seconds 1 --distinct 32 > gpurun_out/c29_bench_nr.json 2> gpurun_out/c29_bench_nr.err
python - <<'PY'
import json
for l in open('gpurun_out/c29_bench_nr.json'):
    if l.startswith('{'):
        d=json.loads(l); print('NORESTART default', d['ms_per_step'], d['roofline']['kernel_ms'], d['config']['parity']['max_abs_rgb_diff_vs_oracle'])
PY
