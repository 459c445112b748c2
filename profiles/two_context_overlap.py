"""Do K1 (integer pipe) and K2 (FP pipe) of two half batches overlap when they are launched on two contexts / streams?
One context decodes 1024 resident frames per step; two contexts decode 512 each, launched alternately, so that one
half's K2 can run beside the other half's K0/K1 wherever the SMs have room for both.
usage (GPU box): python profiles/two_context_overlap.py [batch [distinct [steps]]]"""
import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch
import jpeglibrary_b200 as J, synth

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
ndist = int(sys.argv[2]) if len(sys.argv) > 2 else 32
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
distinct = [synth.encode_jpeg(synth.synth_rgb(1000 + i, 3840, 2160), quality=85, subsampling="4:2:0", restart_rows=1) for i in range(ndist)]
blobs = [distinct[i % ndist] for i in range(batch)]
ctx = [J.Context(0), J.Context(0)]
streams = [torch.cuda.ExternalStream(c.stream, device=0) for c in ctx]


def timed(decs, label):
    for d in decs:
        d.upload()
    for _ in range(3):
        for d in decs:
            d.launch()
    for c in ctx:
        c.synchronize()
    used = streams[:len(decs)]
    e0 = [torch.cuda.Event(enable_timing=True) for _ in used]
    e1 = [torch.cuda.Event(enable_timing=True) for _ in used]
    for e, s in zip(e0, used):
        e.record(s)
    for _ in range(steps):
        for d in decs:
            d.launch()
    for e, s in zip(e1, used):
        e.record(s)
    for c in ctx:
        c.synchronize()
    ms = max(a.elapsed_time(b) for a in e0 for b in e1) / steps
    for d in decs:
        d.finish()
        assert d.status() == [0] * d.count if hasattr(d, "count") else True
    print("%-44s %.3f ms per %d frames" % (label, ms, batch))
    return ms


one = J.JpegBatchDecoder(blobs, J.JB_OUT_RGB24, context=ctx[0], device_output=True)
a = timed([one], "one context, %d frames per launch" % batch)
one.close()
for parts in (2, 4):
    n = batch // parts
    decs = [J.JpegBatchDecoder(blobs[i * n:(i + 1) * n], J.JB_OUT_RGB24, context=ctx[i % 2], device_output=True) for i in range(parts)]
    b = timed(decs, "two contexts, %d launches of %d frames" % (parts, n))
    for d in decs:
        d.close()
