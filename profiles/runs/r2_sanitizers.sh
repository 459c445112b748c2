# compute-sanitizer racecheck / synccheck / memcheck over the GPU tests that exercise every kernel with small inputs
# (the hand-off of K1, the warp-cooperative refinement decoder of K1c, the encoder's packing).  Records -> profiles/.
set -x
mkdir -p gpurun_out
SEL='synthetic_streams or self_synchronising or progressive_synthetic or lossless_synthetic or final_code_in_the_padding or eoi_at_a_restart or bit_identical_to_the_oracle or optimize_synthetic or several_scans'
for tool in racecheck synccheck; do
  ( echo "compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests -m gpu -q -x -k '$SEL'"; 
    timeout 2400 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests -m gpu -q -x -k "$SEL" 2>&1 | tail -40 ) > gpurun_out/r2_sanitizer_$tool.txt
  tail -4 gpurun_out/r2_sanitizer_$tool.txt
done
( echo "compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -q -x -k 'fuzz or corrupted or eoi_at or several_scans or lossless_synthetic or large_batch_packed or truncated or final_code or outgrow or padding'";
  timeout 2400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -q -x -k 'fuzz or corrupted or eoi_at or several_scans or lossless_synthetic or large_batch_packed or truncated or final_code or outgrow or padding' 2>&1 | tail -40 ) > gpurun_out/r2_sanitizer_memcheck.txt
tail -4 gpurun_out/r2_sanitizer_memcheck.txt
