set -x
mkdir -p gpurun_out
for l in 16 4 1; do
  JB_K1C_LANES=$l timeout 300 python profiles/prog_trace.py 1024 8 > gpurun_out/c20_trace_l$l.txt 2>&1
  tail -12 gpurun_out/c20_trace_l$l.txt
done
JB_K1C_LANES=2 timeout 300 python profiles/prog_trace.py 256 8 > gpurun_out/c20_trace_b256_l2.txt 2>&1
tail -12 gpurun_out/c20_trace_b256_l2.txt
