"""K1c schedule of a progressive batch (jb_decode_batch_scan_trace): per scan start / end / time spent waiting for
producer scans.  usage (on a GPU box): python profiles/prog_trace.py [batch [distinct images]]"""
import sys, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import jpeglibrary_b200 as J, synth
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ndist = int(sys.argv[2]) if len(sys.argv) > 2 else 4
distinct = [synth.encode_jpeg(synth.synth_rgb(i, 1920, 1080), quality=85, subsampling="4:4:4", progressive=True) for i in range(ndist)]
blobs = [distinct[i % ndist] for i in range(batch)]
with J.JpegBatchDecoder(blobs, J.JB_OUT_RGB24, device_output=True) as b:
    b.run()
    b.set_profiling(True, trace=True)
    b.upload(); b.launch(); b.finish()
    print(b.profile())
    tr = b.scan_trace()
    t0 = min(t[3] for t in tr)
    print("jobs", len(tr), "span ms", (max(t[4] for t in tr) - t0) / 1e6)
    for img in (0, batch // 2, batch - 1):
        print("image", img)
        for t in sorted([t for t in tr if t[0] == img], key=lambda t: t[1]):
            print("  scan %2d start %8.2f end %8.2f  busy %8.2f  waited %8.2f ms" % (t[1], (t[3] - t0) / 1e6, (t[4] - t0) / 1e6, (t[4] - t[3] - t[5]) / 1e6, t[5] / 1e6))
    # per scan averages
    for s in range(10):
        xs = [t for t in tr if t[1] == s]
        print("scan %d: max end %.1f;" % (s, max((t[4]-t0)/1e6 for t in xs)), end=" ")
        print("mean start %.1f end %.1f busy %.1f waited %.1f" % (np.mean([(t[3]-t0)/1e6 for t in xs]), np.mean([(t[4]-t0)/1e6 for t in xs]), np.mean([(t[4]-t[3]-t[5])/1e6 for t in xs]), np.mean([t[5]/1e6 for t in xs])))
