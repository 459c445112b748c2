# final compute-sanitizer pass of round 2 (after per-component limits, first-error keys, status clearing, row-wise D2H)
set -x
mkdir -p gpurun_out
bash profiles/runs/r2_sanitizers.sh
ls -la gpurun_out
