# batched H2D submission (cudaMemcpyBatchAsync) A/B; full GPU test suite with the package-merge tests
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/c26_tests.log 2>&1; tail -5 gpurun_out/c26_tests.log
for v in 1 0 1; do
  export JB_BATCH_MEMCPY=$v
  timeout 600 python bench.py --workload restart --steps 5 --warmup 3 --e2e-batch 512 --cpu-seconds 1 --distinct 32 > gpurun_out/c26_bench_$v.json 2> gpurun_out/c26_bench_$v.err
  tail -3 gpurun_out/c26_bench_$v.err
  python - "$v" <<'PY'
import json,sys
for l in open('gpurun_out/c26_bench_%s.json'%sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); h=d['value_h2d']; e=d['e2e']
        print('VARIANT batch_memcpy',sys.argv[1], 'step', d['ms_per_step'], 'h2d step', h['ms_per_step'], 'copies alone', h['copies_alone_ms_per_step'], h['copies_alone_gb_per_s'], 'e2e', e['value'], e.get('fraction_of_bare_d2h'))
PY
done
