set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_decode.py -m gpu -q -x > gpurun_out/c18_tests.log 2>&1; tail -3 gpurun_out/c18_tests.log
for v in "" k2o5 k2mad ""; do
  if [ -n "$v" ]; then export JB_LIBDIR=$PWD/jpeglibrary_b200/lib_$v; else unset JB_LIBDIR; fi
  timeout 600 python bench.py --workload restart --steps 5 --warmup 3 --e2e-batch 32 --cpu-seconds 1 --distinct 32 > gpurun_out/c18_bench_$v.json 2> gpurun_out/c18_bench_$v.err
  python - "$v" <<'PY'
import json,sys
for l in open('gpurun_out/c18_bench_%s.json'%sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print('VARIANT',sys.argv[1] or 'default', d['ms_per_step'], d['roofline']['kernel_ms'], d['config']['parity']['max_abs_rgb_diff_vs_oracle'])
PY
done
unset JB_LIBDIR
