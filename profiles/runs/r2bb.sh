set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/c48_tests.log 2>&1; tail -5 gpurun_out/c48_tests.log
timeout 600 python bench.py --workload norestart --steps 5 --warmup 3 --e2e-batch 32 --cpu-seconds 1 --distinct 32 > gpurun_out/c48_bench_nr.json 2> gpurun_out/c48_bench_nr.err
python - <<'PY'
import json
for l in open('gpurun_out/c48_bench_nr.json'):
    if l.startswith('{'):
        d=json.loads(l); print('NORESTART', d['ms_per_step'], d['roofline']['kernel_ms'], d['config']['parity']['max_abs_rgb_diff_vs_oracle'])
PY
