# K1c: relaxed progress polls (no L1 invalidation), thin refinement A/B
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_decode.py -m gpu -q -x -k "progressive or golden or sequential_frames" > gpurun_out/c22_tests.log 2>&1; tail -4 gpurun_out/c22_tests.log
for v in thin coop lanes4 lanes8 lanes32; do
  unset JB_K1C_THIN JB_K1C_LANES
  case $v in coop) export JB_K1C_THIN=0;; lanes4) export JB_K1C_LANES=4;; lanes8) export JB_K1C_LANES=8;; lanes32) export JB_K1C_LANES=32;; esac
  timeout 600 python bench.py --workload progressive --steps 3 --warmup 3 --e2e-batch 32 --cpu-seconds 1 --distinct 32 > gpurun_out/c22_bench_$v.json 2> gpurun_out/c22_bench_$v.err
  tail -3 gpurun_out/c22_bench_$v.err
  python - "$v" <<'PY'
import json,sys
for l in open('gpurun_out/c22_bench_%s.json'%sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print('VARIANT',sys.argv[1], d['ms_per_step'], d['roofline']['kernel_ms'], d['config'].get('parity'))
PY
done
unset JB_K1C_THIN JB_K1C_LANES
timeout 300 python profiles/prog_trace.py 1024 8 > gpurun_out/c22_trace.txt 2>&1
tail -11 gpurun_out/c22_trace.txt
