# K1c whole-warp refinement decoder: lanes per packed warp (the AC first / DC scans) A/B
set -x
mkdir -p gpurun_out
export JB_K1C_THIN=0
for v in 2 4 8 16; do
  export JB_K1C_LANES=$v
  timeout 600 python bench.py --workload progressive --steps 3 --warmup 3 --e2e-batch 32 --cpu-seconds 1 --distinct 32 > gpurun_out/c23_bench_$v.json 2> gpurun_out/c23_bench_$v.err
  tail -3 gpurun_out/c23_bench_$v.err
  python - "$v" <<'PY'
import json,sys
for l in open('gpurun_out/c23_bench_%s.json'%sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print('VARIANT coop lanes',sys.argv[1], d['ms_per_step'], d['roofline']['kernel_ms'], d['config'].get('parity'))
PY
done
JB_K1C_LANES=4 timeout 300 python profiles/prog_trace.py 1024 8 > gpurun_out/c23_trace_l4.txt 2>&1
tail -11 gpurun_out/c23_trace_l4.txt
unset JB_K1C_LANES
timeout 300 python profiles/prog_trace.py 1024 8 > gpurun_out/c23_trace_l32.txt 2>&1
tail -11 gpurun_out/c23_trace_l32.txt
