# K1c: did the per-component limits / first-error reporting cost time?  lib_prev = commit f64d793 (before both)
set -x
mkdir -p gpurun_out
for v in prev "" prev ""; do
  if [ -n "$v" ]; then export JB_LIBDIR=$PWD/jpeglibrary_b200/lib_$v; else unset JB_LIBDIR; fi
  timeout 600 python bench.py --workload progressive --steps 3 --warmup 3 --e2e-batch 32 --cpu-seconds 1 --distinct 32 > gpurun_out/c43_bench_$v.json 2> gpurun_out/c43_bench_$v.err
  tail -2 gpurun_out/c43_bench_$v.err
  python - "$v" <<'PY'
import json,sys
for l in open('gpurun_out/c43_bench_%s.json'%sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print('VARIANT',sys.argv[1] or 'current', d['ms_per_step'], d['roofline']['kernel_ms'])
PY
done
