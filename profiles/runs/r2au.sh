set -x
mkdir -p gpurun_out
for i in 1 2; do
timeout 600 python bench.py --workload progressive --steps 3 --warmup 3 --e2e-batch 32 --cpu-seconds 1 > gpurun_out/c44_bench_$i.json 2> gpurun_out/c44_bench_$i.err
python - "$i" <<'PY'
import json,sys
for l in open('gpurun_out/c44_bench_%s.json'%sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print('RUN',sys.argv[1], d['ms_per_step'], d['roofline']['kernel_ms'], d['config']['distinct_images'])
PY
done
