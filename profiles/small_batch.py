"""Kernel time of small batches of 4K frames (restart-coded, plain, progressive): where batch-level parallelism runs out.
usage (on a GPU box): python profiles/small_batch.py"""
import sys, time, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import jpeglibrary_b200 as J, synth
rgb = [synth.synth_rgb(i, 3840, 2160) for i in range(4)]
kinds = {"restart": dict(quality=85, subsampling="4:2:0", restart_rows=1), "plain": dict(quality=85, subsampling="4:2:0"),
         "progressive": dict(quality=85, subsampling="4:2:0", progressive=True)}
for name, kw in kinds.items():
    blobs4 = [synth.encode_jpeg(r, **kw) for r in rgb]
    for n in (1, 4, 16, 64):
        blobs = [blobs4[i % 4] for i in range(n)]
        with J.JpegBatchDecoder(blobs, J.JB_OUT_RGB24, device_output=True) as b:
            b.run(); b.set_profiling(True)
            for _ in range(5):
                b.upload(); b.launch(); b.finish()
            prof = b.profile()
        tot = sum(ms for _, ms in prof)
        print(f"{name:12s} batch {n:3d}: kernels {tot:7.2f} ms  ({8.2944 * n / tot:7.1f} GP/s)  " + ", ".join(f"{k.replace('jb_','')} {ms:.2f}" for k, ms in prof))

# wall time of the single-image public API (JpegDecoder.Decode into pinned host RGB): marker walk + plan + H2D + kernels + D2H
ctx = J.Context(0)
for name, kw in kinds.items():
    if name == "progressive":
        continue
    blob = synth.encode_jpeg(rgb[0], **kw)
    pinned_in = ctx.pinned_array(len(blob)); pinned_in[:] = np.frombuffer(blob, np.uint8)
    out = ctx.pinned_array(2160 * 3840 * 3).reshape(2160, 3840, 3)
    def once():
        dec = J.JpegDecoder(ctx)
        dec.SetInput(pinned_in)
        dec.SetOutputWriter(J.CudaOutputWriter(out))
        dec.Decode()
    for _ in range(3):
        once()
    t0 = time.perf_counter()
    for _ in range(20):
        once()
    dt = (time.perf_counter() - t0) / 20
    print(f"{name:12s} JpegDecoder.Decode() wall {dt * 1e3:6.2f} ms per 4K frame ({8.2944 / dt / 1e3:5.2f} GP/s)")
