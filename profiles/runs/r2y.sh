# K1c: consumers resume JB_K1C_LAG units behind their producer; sign bit leaves with the code; branch-free placement
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_decode.py -m gpu -q -x -k "progressive or golden or sequential_frames" > gpurun_out/c25_tests.log 2>&1; tail -4 gpurun_out/c25_tests.log
timeout 600 python -m pytest tests/test_gpu_fuzz.py -m gpu -q -x > gpurun_out/c25_fuzz.log 2>&1; tail -3 gpurun_out/c25_fuzz.log
for v in "" lag0 lag32 ""; do
  if [ -n "$v" ]; then export JB_LIBDIR=$PWD/jpeglibrary_b200/lib_$v; else unset JB_LIBDIR; fi
  timeout 600 python bench.py --workload progressive --steps 3 --warmup 3 --e2e-batch 32 --cpu-seconds 1 --distinct 32 > gpurun_out/c25_bench_$v.json 2> gpurun_out/c25_bench_$v.err
  tail -3 gpurun_out/c25_bench_$v.err
  python - "$v" <<'PY'
import json,sys
for l in open('gpurun_out/c25_bench_%s.json'%sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print('VARIANT',sys.argv[1] or 'default(lag96)', d['ms_per_step'], d['roofline']['kernel_ms'], d['config'].get('parity'))
PY
done
unset JB_LIBDIR
