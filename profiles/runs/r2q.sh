set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/c17_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/c17_tests.log
tail -15 gpurun_out/c17_tests.log
timeout 300 python -m pytest tests/test_gpu_fuzz.py -m gpu -q -s -k render_the_same 2>&1 | grep -i "identical\|passed\|failed" | head
timeout 600 python bench.py --workload restart --steps 5 --warmup 3 --e2e-batch 64 --cpu-seconds 1 > gpurun_out/c17_bench.json 2> gpurun_out/c17_bench.err
python - <<'PY'
import json
for l in open('gpurun_out/c17_bench.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['ms_per_step'], d['roofline']['kernel_ms'], d['config']['parity'])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'jb_k2_idct_color_warp' -s 2 -c 1 -o gpurun_out/c17_prof python bench.py --workload restart --distinct 16 --steps 1 --warmup 2 --e2e-batch 32 --cpu-seconds 1 > gpurun_out/c17_ncu.log 2>&1
tail -2 gpurun_out/c17_ncu.log | cut -c1-200
