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
