set -x
mkdir -p gpurun_out
for c in 2 3 4 2 3; do
  JB_E2E_CONTEXTS=$c timeout 600 python bench.py --workload restart --steps 5 --warmup 3 --e2e-batch 512 --cpu-seconds 1 --distinct 32 > gpurun_out/c50_bench_$c.json 2> gpurun_out/c50_bench_$c.err
  python - "$c" <<'PY'
import json,sys
for l in open('gpurun_out/c50_bench_%s.json'%sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); e=d['e2e']; print('CONTEXTS',sys.argv[1], 'e2e', e['value'], e['d2h_gb_per_s_all_ranks'], e['bare_pinned_d2h_gb_per_s_all_ranks'], e['fraction_of_bare_d2h'])
PY
done
